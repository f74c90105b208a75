"""
utils.py -- the two signal helpers of exptool.utils.utils that the coefficient time-series consumers use
(eof.calculate_eof_phase).  Host NumPy, as in the reference: a few hundred numbers per series.
"""
from math import factorial

import numpy as np


def savitzky_golay(y, window_size, order, deriv=0, rate=1):
    '''
    utils.savitzky_golay (utils.py:389-472): least-squares polynomial smoothing (or its deriv-th derivative) over an
    odd window, the ends padded by reflecting the signal about its end values.  (The reference builds the design
    matrix with np.mat, which NumPy 2 removed; the same pseudo-inverse row is taken from a plain array here.)
    '''
    try:
        window_size = abs(int(window_size))
        order = abs(int(order))
    except (TypeError, ValueError):
        raise ValueError("window_size and order have to be of type int")
    if window_size % 2 != 1 or window_size < 1:
        raise TypeError("window_size size must be a positive odd number")
    if window_size < order + 2:
        raise TypeError("window_size is too small for the polynomials order")
    y = np.asarray(y, dtype=np.float64)
    half = (window_size - 1) // 2
    k = np.arange(-half, half + 1, dtype=np.float64)
    design = k[:, None] ** np.arange(order + 1)[None, :]
    taps = np.linalg.pinv(design)[deriv] * rate ** deriv * factorial(deriv)
    head = y[0] - np.abs(y[1:half + 1][::-1] - y[0])
    tail = y[-1] + np.abs(y[-half - 1:-1][::-1] - y[-1])
    return np.convolve(taps[::-1], np.concatenate((head, y, tail)), mode='valid')


def unwrap_phase(times, phases, max_periods=1000):
    '''
    utils.unwrap_phase (utils.py:474-525): every sample is moved by the multiple of 2 pi (within +-max_periods)
    that brings it closest to the previous unwrapped sample, so the series may wind either way.
    As in the reference, a series with negative values is shifted by +pi IN PLACE first (the caller's array changes).
    '''
    times = np.asarray(times)
    if times.size != phases.size:
        raise ValueError('unwrap_phases: times and phases must be equal size.')
    if (np.nanmax(phases) > 2. * np.pi) | (np.nanmin(phases) < -np.pi):
        raise ValueError('unwrap_phases: Values are outside of the accepted phase boundaries.')
    if np.nanmin(phases) < 0:
        phases += np.pi
    if np.nanmax(phases) - np.nanmin(phases) < np.pi:
        print('unwrap_phases WARNING: Were the phases calculated with arctan2?')
    out = np.zeros(phases.size)
    out[0] = phases[0]
    twopi = 2 * np.pi
    for t in range(1, times.size):
        k = np.round((out[t - 1] - phases[t]) / twopi)
        if not np.isfinite(k):
            # a NaN neighbour: the reference's argmin over NaNs picks the first candidate (offset -max_periods)
            k = -max_periods
        k = min(max(k, -max_periods), max_periods - 1)
        # candidates k-1, k, k+1 settle rounding at the half-way point the way argmin does (first minimum)
        cand = phases[t] + twopi * np.array([k - 1, k, k + 1])
        cand = cand[(np.array([k - 1, k, k + 1]) >= -max_periods) & (np.array([k - 1, k, k + 1]) <= max_periods - 1)]
        out[t] = cand[np.abs(cand - out[t - 1]).argmin()] if np.isfinite(out[t - 1]) and np.isfinite(phases[t]) \
            else phases[t] + twopi * (-max_periods)
    return out
