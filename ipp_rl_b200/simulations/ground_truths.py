"""Ground-truth field generators (reset path; reference simulations/ground_truths.py:7-33).

Host-side NumPy, vectorised (the reference fills the spectral amplitude with a Python double loop:
0.25 s per 400x400 map).  Draws the same ``np.random`` numbers in the same order as the reference, so
a seed gives the same field (to rounding of the vectorised power)."""
import numpy as np


def fft_indices(n: int) -> np.ndarray:
    """Integer wave numbers in FFT order as the reference builds them (:7-11): 0..n//2 followed by
    -(n//2 - 1)..-1.  (For odd n this list has n-1 entries; the remaining spectrum row/column keeps
    amplitude 0, as in the reference.)"""
    return np.concatenate([np.arange(0, n // 2 + 1), -np.arange(n // 2 - 1, 0, -1)])


def gaussian_random_field(pk, x_dim: int, y_dim: int) -> np.ndarray:
    """2-D Gaussian random field with power spectrum ``pk(|k|)``, min-max normalised to [0, 1]:
    white noise -> FFT -> times sqrt(pk(|k|)) (0 at k = 0) -> inverse FFT, real part."""
    noise = np.fft.fft2(np.random.normal(size=(y_dim, x_dim)))
    ky, kx = fft_indices(y_dim), fft_indices(x_dim)
    amplitude = np.zeros((y_dim, x_dim))
    k = np.sqrt(ky[:, None].astype(np.float64) ** 2 + kx[None, :].astype(np.float64) ** 2)
    with np.errstate(divide="ignore", invalid="ignore"):
        amp = np.sqrt(pk(k))
    amp[0, 0] = 0.0
    amplitude[: len(ky), : len(kx)] = amp
    field = np.fft.ifft2(noise * amplitude).real
    return (field - field.min()) / (field.max() - field.min())
