"""TEST INFRASTRUCTURE — CPU restatement (numpy float64) of the reference's per-utterance feature extraction and batch
collation, SURVEY.md §8f row N3.  Not product code.

  reference deepspeech_pytorch/loader/data_loader.py
    :73-94    AudioParser.compute_spectrogram -> compute_spectrogram
    :247-270  _collate_fn                     -> collate

The STFT itself lives in the third-party dependency `librosa` (requirements.txt:4, UNPINNED; absent from this image,
so the reference's own output cannot be generated here: PARITY UNPINNED against librosa for this row).  What is restated
is librosa's published algorithm for `librosa.stft(y, n_fft, hop_length, win_length=n_fft, window, center=True)`:
pad n_fft//2 samples on both sides (`pad_mode`: "constant" zeros since librosa 0.10, "reflect" before), frame with
stride hop (1 + len(y)//hop frames), multiply by scipy.signal.get_window(window, n_fft, fftbins=True), rfft; then
`librosa.magphase` -> |D|, `np.log1p`, and torch's `(x - mean) / std` with the unbiased std.  tests/test_spect.py pins
this file against the independent `scipy.signal.stft` and `torch.stft`.
"""
import numpy as np


def get_window(name: str, n: int) -> np.ndarray:
    import scipy.signal
    return scipy.signal.get_window(name, n, fftbins=True)


def stft_mag(y: np.ndarray, n_fft: int, hop: int, window: str = "hamming", pad_mode: str = "constant") -> np.ndarray:
    y = np.asarray(y, dtype=np.float64)
    yp = np.pad(y, n_fft // 2, mode=pad_mode)
    n_frames = 1 + len(y) // hop
    w = get_window(window, n_fft)
    frames = np.stack([yp[f * hop:f * hop + n_fft] for f in range(n_frames)], 0) * w[None, :]
    return np.abs(np.fft.rfft(frames, axis=1)).T          # (n_fft/2+1, n_frames)


def compute_spectrogram(y, sample_rate=16000, window_size=0.02, window_stride=0.01, window="hamming", normalize=True,
                        pad_mode="constant") -> np.ndarray:
    n_fft = int(sample_rate * window_size)
    hop = int(sample_rate * window_stride)
    spect = np.log1p(stft_mag(y, n_fft, hop, window, pad_mode))
    if normalize:
        spect = (spect - spect.mean()) / spect.std(ddof=1)     # torch.Tensor.std() is unbiased
    return spect


def collate(samples):
    """samples: list of (spect (F,T) ndarray, transcript list[int]) -> the reference's batch tuple as numpy arrays"""
    order = sorted(range(len(samples)), key=lambda i: samples[i][0].shape[1], reverse=True)
    F = samples[0][0].shape[0]
    Tmax = samples[order[0]][0].shape[1]
    inputs = np.zeros((len(samples), 1, F, Tmax), np.float32)
    pct = np.zeros(len(samples), np.float32)
    tsz = np.zeros(len(samples), np.int32)
    targets = []
    for row, i in enumerate(order):
        s, tr = samples[i]
        inputs[row, 0, :, :s.shape[1]] = s
        pct[row] = s.shape[1] / float(Tmax)
        tsz[row] = len(tr)
        targets.extend(tr)
    return inputs, np.asarray(targets, np.int64), pct, tsz
