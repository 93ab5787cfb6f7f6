"""TEST INFRASTRUCTURE -- CPU oracle, never imported by the product path.

Restatement of the Timmer & Koenig (1995) power-law noise synthesis that the reference
obtains from the third-party pip package ``colorednoise`` (function
``powerlaw_psd_gaussian``).  The package is NOT vendored in /root/reference and is NOT
pinned in its environment.yaml, so: **parity unpinned** for this function (no golden
vector exists anywhere in the reference).  Call sites it stands in for:
  src/algorithm/tdmpc_icem_similarity_mlp.py:128,149,152
  src/algorithm/tdmpc_icem_similarity_drnn.py:137,158,161

Algorithm (published): draw independent Gaussian real/imag Fourier coefficients with
standard deviation f^(-beta/2) on the rfft grid of the last axis, make the DC and (for
even n) Nyquist bins real, inverse-rfft, and divide by the analytic standard deviation
so each sample has (approximately) unit variance.

Because every sample is a fixed linear map of white Gaussians, ``synthesis_matrix``
below returns that map explicitly; the CUDA sampler multiplies white noise by the same
matrix, which is how the device generator is validated (tests/test_oracle_golden.py::test_colored_noise_matrix_is_the_generator,
tests/test_gpu_parity.py::test_icem_builtin_colored_sampler_statistics).
"""
import numpy as np


def _spectrum_scale(beta, n, fmin=0.0):
    freqs = np.fft.rfftfreq(n)
    fmin = max(fmin, 1.0 / n)
    scale = freqs.copy()
    below = int(np.sum(scale < fmin))
    if 0 < below < len(scale):
        scale[:below] = scale[below]
    scale = scale ** (-beta / 2.0)
    # analytic std of the synthesised series
    w = scale[1:].copy()
    w[-1] *= (1 + (n % 2)) / 2.0
    sigma = 2.0 * np.sqrt(np.sum(w ** 2)) / n
    return scale, sigma


def powerlaw_psd_gaussian(exponent, size, fmin=0, random_state=None):
    """Gaussian (1/f)**exponent noise, unit variance, correlated along the LAST axis."""
    rng = np.random if random_state is None else random_state
    shape = [size] if np.isscalar(size) else list(size)
    n = shape[-1]
    scale, sigma = _spectrum_scale(exponent, n, fmin)
    shape[-1] = len(scale)
    bscale = scale.reshape((1,) * (len(shape) - 1) + (-1,))
    re = rng.normal(scale=bscale, size=shape)
    im = rng.normal(scale=bscale, size=shape)
    if n % 2 == 0:            # Nyquist bin is real
        im[..., -1] = 0.0
        re[..., -1] *= np.sqrt(2.0)
    im[..., 0] = 0.0          # DC bin is real
    re[..., 0] *= np.sqrt(2.0)
    return np.fft.irfft(re + 1j * im, n=n, axis=-1) / sigma


def synthesis_matrix(beta, n):
    """[n, n] float64 matrix S with  colored = S @ white  (white ~ N(0, I_n)).

    Column order of the white vector: real parts of bins 0..n//2, then the imaginary
    parts of the bins that keep one (1..ceil(n/2)-1).  diag(S S^T) is a constant slightly above 1: the
    package normalises by the analytic std of the non-DC bins only; reproduced as is.
    """
    scale, sigma = _spectrum_scale(beta, n)
    nb = len(scale)
    cols = []
    for k in range(nb):                      # real parts
        spec = np.zeros(nb, dtype=complex)
        amp = scale[k]
        if k == 0 or (n % 2 == 0 and k == nb - 1):
            amp *= np.sqrt(2.0)
        spec[k] = amp
        cols.append(np.fft.irfft(spec, n=n) / sigma)
    last_im = nb - 1 if n % 2 == 0 else nb
    for k in range(1, last_im):              # imaginary parts
        spec = np.zeros(nb, dtype=complex)
        spec[k] = 1j * scale[k]
        cols.append(np.fft.irfft(spec, n=n) / sigma)
    S = np.stack(cols, axis=1)
    assert S.shape == (n, n), S.shape
    return S
