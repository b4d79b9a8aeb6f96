"""Row N3 (input pipeline).  CPU: the oracle restatement of librosa's STFT is pinned against the independent
scipy.signal.stft, the window generator against scipy's, the collate bookkeeping against the oracle's.  -m gpu: the
CUDA batch kernel (C-ABI ds2_spectrogram_batch) against the oracle on ragged batches."""
import numpy as np
import pytest
import torch

import deepspeech_pytorch_b200 as ds
from deepspeech_pytorch_b200.input_pipeline import SpectrogramBatcher, analysis_window
from oracle import spect_oracle as SO


def _waves(n, seed, lo=2000, hi=9000):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        L = int(rng.integers(lo, hi))
        t = np.arange(L) / 16000.0
        y = 0.3 * np.sin(2 * np.pi * (200 + 37 * i) * t) + 0.05 * rng.standard_normal(L)
        out.append(y.astype(np.float32))
    return out


@pytest.mark.parametrize("pad_mode,boundary", [("constant", "zeros"), ("reflect", "even")])
def test_oracle_stft_matches_scipy(pad_mode, boundary):
    import scipy.signal
    y = _waves(1, 3)[0]
    n_fft, hop = 320, 160
    mag = SO.stft_mag(y, n_fft, hop, "hamming", pad_mode)
    w = scipy.signal.get_window("hamming", n_fft, fftbins=True)
    f, t, Z = scipy.signal.stft(y.astype(np.float64), window=w, nperseg=n_fft, noverlap=n_fft - hop, nfft=n_fft,
                                boundary=boundary, padded=False, return_onesided=True)
    ref = np.abs(Z) * w.sum()                      # scipy scales by 1 / sum(window)
    n = min(ref.shape[1], mag.shape[1])
    assert mag.shape[0] == 161 and n >= mag.shape[1] - 1
    assert np.abs(mag[:, :n] - ref[:, :n]).max() <= 1e-9 * max(1.0, ref.max())
    assert mag.shape[1] == 1 + len(y) // hop       # librosa's frame count


@pytest.mark.parametrize("pad_mode", ["constant", "reflect"])
def test_oracle_stft_matches_torch_stft(pad_mode):
    """second independent pin: torch.stft(center=True) documents itself as librosa-compatible (same centring, same frame
    count 1 + len // hop, periodic window) — the whole feature (|STFT| -> log1p -> normalise) in float64"""
    import torch
    y = _waves(1, 5)[0]
    n_fft, hop = 320, 160
    w = torch.hamming_window(n_fft, periodic=True, dtype=torch.float64)
    Z = torch.stft(torch.from_numpy(y).double(), n_fft, hop_length=hop, win_length=n_fft, window=w, center=True,
                   pad_mode=pad_mode, return_complex=True)
    assert Z.shape == (161, 1 + len(y) // hop)
    mag = SO.stft_mag(y, n_fft, hop, "hamming", pad_mode)
    assert np.abs(mag - Z.abs().numpy()).max() <= 1e-9 * max(1.0, float(Z.abs().max()))
    feat = torch.log1p(Z.abs())
    feat = (feat - feat.mean()) / feat.std()
    got = SO.compute_spectrogram(y, pad_mode=pad_mode)
    assert np.abs(got - feat.numpy()).max() <= 1e-9


@pytest.mark.parametrize("name", ["hamming", "hann", "blackman", "bartlett"])
def test_window_generator_matches_scipy(name):
    assert np.abs(analysis_window(name, 320) - SO.get_window(name, 320)).max() < 1e-7


def test_collate_order_is_the_stable_descending_sort():
    lens = [3200, 4800, 3200, 8000, 4800 + 159]
    order, frames = SpectrogramBatcher.order_and_frames(lens, 160)
    assert frames == [21, 31, 21, 51, 31] and order == [3, 1, 4, 0, 2]


def test_no_cpu_path():
    with pytest.raises(ds.Ds2Error):
        SpectrogramBatcher(ds.SpectConfig(), device="cpu")


@pytest.mark.gpu
@pytest.mark.parametrize("pad_mode,normalize", [("constant", True), ("reflect", True), ("constant", False)])
def test_gpu_batch_matches_oracle(pad_mode, normalize):
    waves = _waves(7, 11) + [np.zeros(1700, np.float32) + 0.01]
    rng = np.random.default_rng(5)
    transcripts = [list(rng.integers(1, 29, int(rng.integers(1, 12)))) for _ in waves]
    batcher = SpectrogramBatcher(ds.SpectConfig(), normalize=normalize, pad_mode=pad_mode)
    inputs, targets, pct, tsz = batcher(waves, transcripts)
    torch.cuda.synchronize()
    samples = [(SO.compute_spectrogram(w, normalize=normalize, pad_mode=pad_mode), tr) for w, tr in zip(waves, transcripts)]
    r_in, r_t, r_pct, r_tsz = SO.collate(samples)
    assert tuple(inputs.shape) == r_in.shape and inputs.dtype == torch.float32 and inputs.is_cuda
    assert targets.dtype == torch.int64 and targets.tolist() == r_t.tolist()          # integers: bit-exact
    assert tsz.dtype == torch.int32 and tsz.tolist() == r_tsz.tolist()
    assert pct.dtype == torch.float32 and np.array_equal(pct.numpy(), r_pct)
    got = inputs.cpu().numpy()
    err = np.abs(got - r_in).max()
    print(f"\n[spect] pad={pad_mode} normalize={normalize}: max abs err {err:.2e} (values up to {np.abs(r_in).max():.2f})")
    assert err < 2e-4
    for row in range(len(waves)):                                                        # padding exactly zero
        n = int(round(float(pct[row]) * inputs.shape[3]))
        assert float(np.abs(got[row, 0, :, n:]).max() if n < inputs.shape[3] else 0.0) == 0.0


@pytest.mark.gpu
def test_gpu_batch_feeds_the_train_step():
    from gpu_helpers import make_model
    waves = _waves(4, 2, lo=6000, hi=12000)
    transcripts = [[3, 5, 7], [2, 2, 9, 1], [4], [8, 6]]
    batch = SpectrogramBatcher(ds.SpectConfig())(waves, transcripts)
    ds.set_precision("fp32")
    model = make_model("gru", True, 16, 1).train()
    loss = model.training_step(batch, 0)
    loss.backward()
    assert torch.isfinite(loss) and float(loss) > 0
