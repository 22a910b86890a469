"""CPU check of the algorithm behind b200fno_eval_metrics (csrc/metrics.cu): the radial-bin spectra of
metrics.py:70-104 computed from a TRUNCATED separable DFT (wavenumbers below nb = min(t,h,w)//2 per axis, fp32
twiddle tables indexed by (k*n) mod N, W then H then T) and exact-integer radial bins, against the oracle's fftn."""
import numpy as np
import pytest
import torch

from oracle import metrics_oracle as M


def truncated_spectrum(x, nb):
    """x: [b,t,h,w,c] float64 -> sum |F|^2 per radial bin [b, nb, c], the way the CUDA kernels compute it."""
    b, t, h, w, c = x.shape
    tw = lambda N: np.exp(-2j * np.pi * np.arange(N) / N).astype(np.complex64).astype(np.complex128)
    dft = lambda N: tw(N)[(np.arange(nb)[:, None] * np.arange(N)[None, :]) % N]  # [nb, N], table lookups
    a = np.einsum("kw,bthwc->bthkc", dft(w), x)          # dftw_kernel
    a = np.einsum("jh,bthkc->btjkc", dft(h), a)          # dft_axis_kernel along H
    a = np.einsum("it,btjkc->bijkc", dft(t), a)          # dft_axis_kernel along T
    out = np.zeros((b, nb, c))
    for i in range(nb):
        for j in range(nb):
            for k in range(nb):
                s2 = i * i + j * j + k * k
                it = int(np.floor(np.sqrt(float(s2))))
                while it * it > s2:
                    it -= 1
                while (it + 1) * (it + 1) <= s2:
                    it += 1
                if it <= nb - 1:
                    out[:, it] += np.abs(a[:, i, j, k]) ** 2   # bin_kernel
    return out


@pytest.mark.parametrize("shape", [(2, 8, 10, 12, 3), (3, 9, 7, 11, 2), (2, 12, 6, 16, 1)])
def test_truncated_dft_spectrum_equals_fftn_binning(shape):
    torch.manual_seed(4)
    x = torch.randn(*shape, dtype=torch.float64)
    b, t, h, w, c = shape
    nb = min(t // 2, h // 2, w // 2)
    ref = M._spectrum(torch.fft.fftn(x, dim=[1, 2, 3]), M.radial_bins(t, h, w), nb).numpy()
    got = truncated_spectrum(x.numpy(), nb)
    assert np.allclose(got, ref, rtol=2e-6, atol=1e-9)
