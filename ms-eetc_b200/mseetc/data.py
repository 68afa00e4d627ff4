"""Measured motor + converter losses (two converter configurations), per motor [W], on a grid of load [%] and
stator frequency [Hz].  Same accessor as the reference's ``mseetc/data.py`` (``dataLosses``); the numbers live in
``motor_losses.json`` next to this file."""
import json
import os

_FILE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'motor_losses.json')


def dataLosses():
    with open(_FILE) as fh:
        raw = json.load(fh)
    cfg = lambda key: {'loads': list(raw['loads_percent']), 'frequencies': list(raw['frequencies_hz']),
                       'losses': [list(row) for row in raw[key]]}
    return cfg('config_a_losses_w'), cfg('config_b_losses_w')
