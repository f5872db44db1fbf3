"""CPU restatement of the reference's audio feature front end (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

Follows utils/data_loader.py:60-91 (`SpectrogramParser.parse_audio`) and the batch layout of `_collate_fn` (:182-214):

    n_fft = int(sample_rate * window_size) = 320, hop = int(sample_rate * window_stride) = 160, win_length = n_fft
    D = librosa.stft(y, n_fft, hop_length, win_length, window=scipy.signal.hamming)   # the CALLABLE (data_loader.py:20,52)
    spect = log1p(|D|);  spect = (spect - spect.mean()) / spect.std()          # torch: unbiased std
    inputs[b, 0, :, :T_b] = spect_b  (zero padded to the longest utterance);  input_sizes[b] = T_b

Third-party arithmetic: librosa (not vendored in /root/reference and not installed here; `requirement.sh:1` installs it
unpinned).  What is restated is librosa.stft's published algorithm: window = librosa.filters.get_window(window, n_fft),
which for a callable returns window(n_fft) = scipy.signal.hamming(n_fft) with scipy's default sym=True -- the SYMMETRIC
Hamming window 0.54 - 0.46 cos(2 pi n / (N-1)) (a 'hamming' string would give the periodic one, selectable here); centre padding of n_fft // 2 samples on both sides -- `pad_mode='reflect'` was the default of
the librosa releases contemporary with the reference (<= 0.9; 0.10 switched to zeros, selectable here) --, frames every
`hop` samples, rfft of each windowed frame, 1 + len(y) // hop frames.

PARITY STATUS: unpinned against librosa itself (absent); `tests/test_features.py` pins the window against
scipy.signal.windows.hamming(N, sym=True) (scipy is installed) and the transform against an independent implementation
of the same definition, torch.stft (CPU), to 1e-5.
"""
import numpy as np


def hamming(n_fft: int, periodic: bool = False) -> np.ndarray:
    """Symmetric (reference: callable scipy.signal.hamming -> sym=True) or periodic Hamming window."""
    n = np.arange(n_fft, dtype=np.float64)
    return 0.54 - 0.46 * np.cos(2.0 * np.pi * n / (n_fft if periodic else n_fft - 1))


def stft_log_spectrogram(y: np.ndarray, n_fft: int = 320, hop: int = 160, reflect: bool = True, normalize: bool = True,
                         periodic: bool = False) -> np.ndarray:
    """One utterance -> [n_fft // 2 + 1, 1 + len(y) // hop] float32 (data_loader.py:74-89)."""
    y = np.asarray(y, dtype=np.float64)
    half = n_fft // 2
    yp = np.pad(y, (half, half), mode="reflect" if reflect else "constant")
    n_frames = 1 + len(y) // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(n_frames)[:, None]
    frames = yp[idx] * hamming(n_fft, periodic)[None, :]
    spec = np.fft.rfft(frames, axis=1).T                                   # [bins, frames]
    out = np.log1p(np.abs(spec)).astype(np.float32)
    if normalize:
        x = out.astype(np.float64)
        out = ((x - x.mean()) / x.std(ddof=1)).astype(np.float32)          # torch.Tensor.std is unbiased
    return out


def collate(waves, n_fft: int = 320, hop: int = 160, reflect: bool = True, normalize: bool = True, periodic: bool = False):
    """List of waveforms -> (inputs [B,1,F,T_max] zero padded, input_sizes [B], input_percentages [B]) as _collate_fn."""
    specs = [stft_log_spectrogram(w, n_fft, hop, reflect, normalize, periodic) for w in waves]
    t_max = max(s.shape[1] for s in specs)
    inputs = np.zeros((len(specs), 1, n_fft // 2 + 1, t_max), dtype=np.float32)
    for b, s in enumerate(specs):
        inputs[b, 0, :, : s.shape[1]] = s
    sizes = np.array([s.shape[1] for s in specs], dtype=np.int32)
    return inputs, sizes, (sizes / float(t_max)).astype(np.float32)
