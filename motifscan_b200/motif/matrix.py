"""PFM -> PPM -> PWM conversion rules of the reference (motifscan/motif/matrix.py:74-171),
vectorised over whole matrices.  The PWM values produced here are what the scan path consumes;
they are 5-decimal fixed point by construction (np.around(..., 5), matrix.py:169)."""
import numpy as np

BASES = "ACGT"


def pfm_to_ppm(pfm, pseudo=0.001):
    """matrix.py:74-98 + :125-146: column-normalise; columns containing a zero get the pseudo
    probability `pseudo / (1 - 4 * pseudo)` added to EVERY row and are renormalised."""
    pfm = np.asarray(pfm)
    if not np.issubdtype(pfm.dtype, np.integer) or np.any(pfm < 0):
        raise ValueError("values in PFM should be non-negative integers")
    tot = pfm.sum(axis=0)
    if np.any(tot == 0):
        raise ValueError("all values of a PFM position are 0")
    if not 0 < pseudo < 0.25:
        raise ValueError("the range of pseudo should be (0, 0.25)")
    ppm = pfm / tot
    zero_cols = np.any(ppm == 0, axis=0)
    ppm[:, zero_cols] += pseudo / (1 - 4 * pseudo)
    return ppm / ppm.sum(axis=0)


def ppm_to_pwm(ppm, bg_freq=None):
    """matrix.py:148-171: natural-log odds against the background, rounded to 5 decimals."""
    if bg_freq is None:
        bg_freq = {b: 0.25 for b in BASES}
    bg = np.asarray([bg_freq[b] for b in BASES], dtype=float).reshape(4, 1)
    return np.around(np.log(np.asarray(ppm) / bg), 5)


def pfm_to_pwm(pfm, bg_freq=None, pseudo=0.001):
    return ppm_to_pwm(pfm_to_ppm(pfm, pseudo), bg_freq)
