"""Power-law ("colored") Gaussian noise as a linear map of white noise.

The reference draws iCEM action noise with `colorednoise.powerlaw_psd_gaussian(beta, size=(n, A, T))`
on the host (tdmpc_icem_similarity_mlp.py:126-157): independent Gaussian Fourier coefficients with
standard deviation f^(-beta/2), inverse real FFT along the last axis, unit variance.  For a fixed
length T that is a fixed T x T matrix applied to T white normals, which is what the CUDA sampler
multiplies its Philox stream by (one row of the matrix per planning step).
"""
import numpy as np


def synthesis_matrix(beta, n):
    """[n, n] float64 S with colored = S @ white, white ~ N(0, I).  (The package normalises by the
    analytic std of the non-DC bins only, so diag(S S^T) is a constant slightly above 1 -- reproduced as is.)"""
    if beta == 0 or n == 1:
        return np.eye(n)
    freqs = np.fft.rfftfreq(n)
    amp = np.where(freqs < 1.0 / n, 1.0 / n, freqs) ** (-beta / 2.0)
    tail = amp[1:].copy()
    tail[-1] *= (1 + (n % 2)) / 2.0
    sigma = 2.0 * np.sqrt(np.sum(tail ** 2)) / n
    nb = len(amp)
    cols = []
    for k in range(nb):                                     # real parts of bins 0 .. n//2
        spec = np.zeros(nb, dtype=complex)
        spec[k] = amp[k] * (np.sqrt(2.0) if (k == 0 or (n % 2 == 0 and k == nb - 1)) else 1.0)
        cols.append(np.fft.irfft(spec, n=n))
    for k in range(1, nb - 1 if n % 2 == 0 else nb):        # imaginary parts of the bins that keep one
        spec = np.zeros(nb, dtype=complex)
        spec[k] = 1j * amp[k]
        cols.append(np.fft.irfft(spec, n=n))
    return np.stack(cols, axis=1) / sigma


def color_matrices(len_mix, horizon, beta_shift, betas=(1.0, 2.5)):
    """The [3, len_mix, len_mix] fp32 block the C ABI expects (tdmpc_plan_args.color_mats): the two colored
    classes of sample_mix_action_sequence (series of length cfg.horizon, truncated to the plan horizon) and
    the tail sampler of the shifted elites (series of length plan horizon, beta = cfg.noise_beta)."""
    out = np.zeros((3, len_mix, len_mix), dtype=np.float32)
    out[0] = synthesis_matrix(betas[0], len_mix)
    out[1] = synthesis_matrix(betas[1], len_mix)
    out[2, :horizon, :horizon] = synthesis_matrix(beta_shift, horizon)
    return out


def mix_split(n):
    """(white, pink, brown) row counts of sample_mix_action_sequence (tdmpc_icem_similarity_mlp.py:139-147)."""
    q, r = divmod(n, 3)
    return q, q + (1 if r >= 1 else 0), q + (1 if r == 2 else 0)
