"""Track description and space discretisation for the B200 solver.

Same public surface as the reference's ``mseetc/track.py`` (``Track`` with ``speedLimits`` / ``gradients`` /
``curvatures`` data frames, ``mergeDataFrames``, ``updateLimits``, ``reverse``, ``sampleClothoid``,
``computeDiscretizationPoints``), re-implemented on numpy step tables (sorted breakpoints + values) so that
it works with current pandas and so that the per-interval SoA tables the kernels need (``ds``, ``c0``,
``bmax``) fall out directly.  Reference behaviour followed: track.py:91-107 (grid), :114-169 (JSON import),
:270-348 (clothoid sampling), :351-374 (reverse), :377-383 (merge), :420-450 (crop).
"""
import json
import sys
from pathlib import Path

import numpy as np
import pandas as pd

from mseetc.utils import checkTTOBenchVersion, convertUnit

_POS = 'Position [m]'


def importTuples(tuples, xLabel, yLabels):
    "List of (position, value, ...) -> data frame indexed by position."
    if not isinstance(yLabels, list):
        yLabels = [yLabels]
    if not isinstance(tuples, list):
        raise ValueError("Input must be a list (of tuples or lists)!")
    width = 1 + len(yLabels)
    if not all(isinstance(row, (tuple, list)) and len(row) == width for row in tuples):
        raise ValueError("Error in list!")
    position = np.array([row[0] for row in tuples])
    if (position < 0).any():
        raise ValueError("Position data cannot be negative!")
    if np.isinf(position).any():
        raise ValueError("Position data cannot be infinite!")
    if (np.diff(position) <= 0).any():
        raise ValueError("Position data must monotonically increase!")
    data = {label: [float(row[1 + j]) for row in tuples] for j, label in enumerate(yLabels)}
    return pd.DataFrame(data, index=pd.Index(position, name=xLabel))


def checkDataFrame(df, trackLength):
    if df.index[0] != 0:
        raise ValueError("Error in '{}': First track section must start at 0 m (beginning of track)!".format(df.columns[0]))
    if df.index[-1] > trackLength:
        raise ValueError("Error in '{}': Last track section must start before {} m (end of track)!".format(df.columns[0], trackLength))
    return True


def computeAltitude(gradients, length, altitudeStart=0):
    "Altitude profile implied by the gradient sections."
    start = gradients.index.values.astype(float)
    slope = (gradients.iloc[:, 0] if isinstance(gradients, pd.DataFrame) else gradients).values.astype(float)
    edges = np.append(start, length)
    altitude = altitudeStart + np.concatenate([[0.0], np.cumsum(np.diff(edges) * slope / 1e3)])
    return pd.DataFrame({'Altitude [m]': altitude}, index=pd.Index(edges, name=gradients.index.name))


def _sample_steps(tables, positions):
    "Evaluate step functions (frame per quantity) at sorted positions with forward-fill semantics."
    cols = {}
    for frame in tables:
        brk = frame.index.values.astype(float)
        idx = np.searchsorted(brk, positions, side='right') - 1
        for name in frame.columns:
            vals = frame[name].values.astype(float)
            cols[name] = np.where(idx >= 0, vals[np.clip(idx, 0, None)], np.nan)
    return cols


def computeDiscretizationPoints(track, numIntervals):
    """Grid of numIntervals+1 positions = uniform points + every section start (reference track.py:91-107)."""
    merged = track.mergeDataFrames()
    uniform = np.linspace(0, track.length, numIntervals + 1 - (len(merged) - 1))
    grid = np.union1d(uniform, merged.index.values.astype(float))
    if len(grid) != numIntervals + 1:
        raise ValueError("Wrong number of computed discretization intervals!")
    cols = _sample_steps([merged], grid)
    return pd.DataFrame({name: cols[name] for name in merged.columns}, index=pd.Index(grid, name='position [m]'))


class Track():

    CURVATURE_THRESHOLD = 1 / 150   # largest admissible |curvature| [1/m]

    def __init__(self, config, pathJSON=Path(__file__).parent.parent / 'tracks'):
        if not isinstance(config, dict):
            raise ValueError("Track configuration should be provided as a dictionary!")
        if 'id' not in config:
            raise ValueError("Track ID must be specified in configuration!")
        with open(Path(pathJSON) / (config['id'] + '.json')) as fh:
            data = json.load(fh)
        checkTTOBenchVersion(data, ['1.1', '1.2', '1.3'])

        stops = data['stops']
        self.length = convertUnit(stops['values'][-1], stops['unit'])
        self.altitude = convertUnit(data['altitude']['value'], data['altitude']['unit']) if 'altitude' in data else 0
        self.title = data['metadata']['id']

        self.importSpeedLimitTuples(data['speed limits']['values'], data['speed limits']['units']['velocity'])
        if 'gradients' in data:
            self.importGradientTuples(data['gradients']['values'], data['gradients']['units']['slope'])
        else:
            self.importGradientTuples([(0.0, 0.0)], 'permil')
        if 'curvatures' in data:
            units = data['curvatures']['units']
            self.importCurvatureTuples(data['curvatures']['values'], units['radius at start'], units['radius at end'],
                                       config.get('clothoidSamplingInterval'))
        else:
            self.importCurvatureTuples([(0.0, "infinity", "infinity")], 'm', 'm', config.get('clothoidSamplingInterval'))

        nStops = len(stops['values'])
        iFrom = config.get('from', 0)
        iTo = config.get('to', nStops - 1)
        if not 0 <= iFrom < nStops - 1:
            raise ValueError("Index of departure is out of bounds!")
        if not iFrom < iTo < nStops:
            raise ValueError("Index of destination is out of bounds!")
        self.updateLimits(convertUnit(stops['values'][iFrom], stops['unit']), convertUnit(stops['values'][iTo], stops['unit']))
        self.checkFields()

    @classmethod
    def fromData(cls, length, speedLimits, gradients=((0.0, 0.0),), curvatures=((0.0, "infinity", "infinity"),), title='synthetic',
                 altitude=0, speedUnit='km/h', clothoidSamplingInterval=None):
        """Additive constructor (no reference analogue): a track from in-memory section lists instead of a TTOBench JSON file.
        speedLimits [(position m, limit)], gradients [(position m, permil)], curvatures [(position m, Rstart m, Rend m)]."""
        self = cls.__new__(cls)
        self.length = float(length)
        self.altitude = altitude
        self.title = title
        self.importSpeedLimitTuples([tuple(x) for x in speedLimits], speedUnit)
        self.importGradientTuples([tuple(x) for x in gradients], 'permil')
        self.importCurvatureTuples([tuple(x) for x in curvatures], 'm', 'm', clothoidSamplingInterval)
        self.checkFields()
        return self

    # ------------------------------------------------------------------ validation
    def lengthOk(self):
        return bool(self.length is not None and self.length > 0 and not np.isinf(self.length))

    def gradientsOk(self):
        return bool(self.gradients.shape[0] > 0 and checkDataFrame(self.gradients, self.length))

    def speedLimitsOk(self):
        return bool(self.speedLimits.shape[0] > 0 and checkDataFrame(self.speedLimits, self.length))

    def curvaturesOk(self):
        if (np.abs(self.curvatures['Curvature [1/m]'].values) > Track.CURVATURE_THRESHOLD).any():
            return False
        return bool(self.curvatures.shape[0] > 0 and checkDataFrame(self.curvatures, self.length))

    def checkFields(self):
        if not self.lengthOk():
            raise ValueError("Track length must be a strictly positive number, not {}!".format(self.length))
        if self.altitude is None or np.isinf(self.altitude):
            raise ValueError("Altitude must be a number, not {}!".format(self.altitude))
        if not self.gradientsOk():
            raise ValueError("Issue with track gradients!")
        if not self.speedLimitsOk():
            raise ValueError("Issue with track speed limits!")
        if not self.curvaturesOk():
            raise ValueError("Issue with track curvatures!")

    # ------------------------------------------------------------------ import
    def importGradientTuples(self, tuples, unit='permil'):
        if not self.lengthOk():
            raise ValueError("Cannot import gradients without a valid track length!")
        if unit not in {'permil'}:
            raise ValueError("Specified gradient unit not supported!")
        self.gradients = importTuples(tuples, _POS, 'Gradient [permil]')
        checkDataFrame(self.gradients, self.length)

    def importSpeedLimitTuples(self, tuples, unit='km/h'):
        if not self.lengthOk():
            raise ValueError("Cannot import speed limits without a valid track length!")
        if unit not in {'km/h', 'm/s'}:
            raise ValueError("Specified speed unit not supported!")
        self.speedLimits = importTuples([(p, convertUnit(v, unit)) for p, v in tuples], _POS, 'Speed limit [m/s]')
        checkDataFrame(self.speedLimits, self.length)

    def importCurvatureTuples(self, tuples, unitRadiusStart='m', unitRadiusEnd='m', clothoidSamplingInterval=None):
        if not self.lengthOk():
            raise ValueError("Cannot import curvature without a valid track length!")
        if unitRadiusStart not in {'m', 'km'} or unitRadiusEnd not in {'m', 'km'}:
            raise ValueError("Specified curvature radius unit not supported!")
        # float("infinity") -> inf, i.e. straight track
        sections = [(p, convertUnit(float(r0), unitRadiusStart), convertUnit(float(r1), unitRadiusEnd)) for p, r0, r1 in tuples]
        self.curvatures = importTuples(self.sampleClothoid(sections, clothoidSamplingInterval), _POS, ['Curvature [1/m]'])
        checkDataFrame(self.curvatures, self.length)

    def sampleClothoid(self, tuples, ds=None):
        """Piecewise-constant approximation of clothoid transition curves (reference track.py:270-348).

        A section (p, Rstart, Rend) with linearly varying curvature K(s) is cut into pieces of length ds;
        each piece gets the mean of K at its two ends, the last piece (length in [ds, 2ds)) runs to the
        section end.  Without ds (or when the section is shorter than ds) the whole section gets the mean
        of its end curvatures.  Returns a list of (position, curvature).
        """
        radii = [sec[j] for sec in tuples for j in (1, 2)]
        if any(r == 0 for r in radii):
            raise ValueError("Curvature radius cannot be 0!")
        if any(sec[0] < 0 for sec in tuples):
            raise ValueError("Positions cannot be negative!")
        if any(a[0] == b[0] for a, b in zip(tuples[:-1], tuples[1:])):
            raise ValueError("Positions must be monotonically increasing")
        if ds is not None and ds <= 0:
            raise ValueError("Discretization step must be greater than zero or None!")

        out = []
        for n, (start, rStart, rEnd) in enumerate(tuples):
            kStart, kEnd = 1 / rStart, 1 / rEnd
            if abs(kStart - kEnd) <= sys.float_info.epsilon:
                out.append((start, kStart))
                continue
            end = tuples[n + 1][0] if n + 1 < len(tuples) else self.length
            pieces = 0 if ds is None else int((end - start) / ds)
            if pieces == 0:
                out.append((start, (kStart + kEnd) / 2))
                continue
            alpha = (end - start) / (kEnd - kStart)     # K(s) = K_start + (s - start)/alpha
            for j in range(pieces):
                kHere = kStart + j * ds / alpha
                mean = (kHere + kEnd) / 2 if j == pieces - 1 else kHere + ds / (2 * alpha)
                out.append((start + j * ds, mean))
        return out

    # ------------------------------------------------------------------ transformations
    def reverse(self):
        "Same track travelled in the opposite direction."
        try:
            self.checkFields()
        except ValueError as e:
            raise ValueError("Track cannot be reversed due to error: {}".format(str(e)))

        def mirrored(df, sign):
            ends = np.append(df.index.values[1:], self.length)
            col = df.columns[0]
            return pd.DataFrame({col: sign * df[col].values[::-1]}, index=pd.Index((self.length - ends)[::-1], name=df.index.name))

        self.gradients = mirrored(self.gradients, -1.0)
        self.speedLimits = mirrored(self.speedLimits, 1.0)
        self.curvatures = mirrored(self.curvatures, -1.0)
        self.title = self.title + ' (reversed)'
        return self

    def mergeDataFrames(self):
        "One row per section start of any quantity; columns curvature, gradient, speed limit (forward filled)."
        frames = [self.curvatures, self.gradients, self.speedLimits]
        pos = np.unique(np.concatenate([f.index.values.astype(float) for f in frames]))
        cols = _sample_steps(frames, pos)
        order = [c for f in frames for c in f.columns]
        return pd.DataFrame({c: cols[c] for c in order}, index=pd.Index(pos, name=_POS))

    def print(self):
        print(self.mergeDataFrames())

    def plot(self, figSize=[12, 6]):
        "Speed limits and altitude profile (needs matplotlib, which is optional for the solver)."
        import matplotlib.pyplot as plt
        fig, axV = plt.subplots(figsize=figSize)
        pos = np.append(self.speedLimits.index.values, self.length) / 1e3
        lim = np.append(self.speedLimits.iloc[:, 0].values, self.speedLimits.iloc[-1, 0]) * 3.6
        axV.step(pos, lim, where='post', color='purple', label='Speed limit')
        axV.set_xlabel('Position [km]'); axV.set_ylabel('Velocity [km/h]'); axV.legend(loc='lower left')
        alt = computeAltitude(self.gradients, self.length)
        axA = axV.twinx()
        axA.plot(alt.index.values / 1e3, alt['Altitude [m]'].values, color='gray', label='Track profile')
        axA.set_ylabel('Altitude [m]'); axA.legend(loc='upper right'); axA.grid(True)
        axA.set_title('Visualization of ' + self.title + ' track')
        plt.show()

    def updateLimits(self, positionStart=None, positionEnd=None, unit='m'):
        "Crop the track to [positionStart, positionEnd] and shift positions so that it starts at 0."
        a = 0 if positionStart is None else positionStart
        b = self.length if positionEnd is None else positionEnd
        if (not 0 <= a < self.length) or (not 0 < b <= self.length):
            raise ValueError("Given positions must be between limits of track!")
        a, b = convertUnit(a, unit), convertUnit(b, unit)

        def cropped(df):
            brk = df.index.values.astype(float)
            pos = np.union1d(brk, [a])
            cols = _sample_steps([df], pos)
            keep = (pos >= a) & (pos <= b)
            newPos = pos[keep] - pos[keep][0]
            return pd.DataFrame({c: cols[c][keep] for c in df.columns}, index=pd.Index(newPos, name=_POS))

        self.length -= a + (self.length - b)
        self.speedLimits = cropped(self.speedLimits)
        self.gradients = cropped(self.gradients)
        self.curvatures = cropped(self.curvatures)


if __name__ == '__main__':
    Track(config={'id': 'CH_StGallen_Wil'}).print()
