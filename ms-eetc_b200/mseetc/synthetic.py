"""Synthetic tracks for the batched benchmarks (BASELINE.json configs 4 and 5; recipe in SURVEY.md section 8d).

Not part of the reference: random gradient / curvature / speed-limit profiles with the statistics of the spec --
gradient sections every 500-3000 m with N(0, 6 permil) clipped to +-25 permil, speed limits from {80,100,120,140} km/h every
5-20 km, curve radius infinite with probability 0.7 else U(300, 3000) m (|kappa| <= 1/150, track.py:112)."""
import numpy as np

from mseetc.track import Track


def random_track(rng, length=None, title='synthetic'):
    L = float(rng.uniform(5e3, 50e3)) if length is None else float(length)
    grads, curvs, pos = [], [], 0.0
    while pos < L:
        grads.append((pos, float(np.clip(rng.normal(0.0, 6.0), -25.0, 25.0))))
        radius = "infinity" if rng.uniform() < 0.7 else float(rng.uniform(300.0, 3000.0)) * (1 if rng.uniform() < 0.5 else -1)
        curvs.append((pos, radius, radius))
        pos += float(rng.uniform(500.0, 3000.0))
    limits, pos = [], 0.0
    while pos < L:
        limits.append((pos, float(rng.choice([80, 100, 120, 140]))))
        pos += float(rng.uniform(5e3, 20e3))
    return Track.fromData(L, limits, grads, curvs, title=title)
