"""Feature front end: the numpy oracle is pinned against torch.stft (an independent implementation of
the librosa.stft definition the reference calls, utils/data_loader.py:74-79); the CUDA path is compared with the oracle."""
import numpy as np
import pytest
import torch

from oracle import features_oracle as FO
from tests.helpers import rel_err


def _waves(lengths, seed=0):
    g = np.random.default_rng(seed)
    out = []
    for n in lengths:
        t = np.arange(n) / 16000.0
        y = 0.3 * np.sin(2 * np.pi * (200 + 50 * g.random()) * t) + 0.1 * g.standard_normal(n) * (1 + np.sin(2 * np.pi * 3 * t))
        out.append(y.astype(np.float32))
    return out


def test_oracle_window_is_the_symmetric_hamming_the_reference_computes():
    """utils/data_loader.py:20 maps 'hamming' to the CALLABLE scipy.signal.hamming and :77-78 hands it to librosa.stft;
    librosa.filters.get_window evaluates a callable as window(n_fft), i.e. scipy's default sym=True."""
    from scipy.signal import windows
    assert np.allclose(FO.hamming(320), windows.hamming(320, sym=True), atol=1e-15)
    assert np.allclose(FO.hamming(320, periodic=True), windows.hamming(320, sym=False), atol=1e-15)
    assert abs(FO.hamming(320)[-1] - 0.08) < 1e-12 and FO.hamming(320, periodic=True)[-1] > 0.08


@pytest.mark.parametrize("reflect", [True, False])
@pytest.mark.parametrize("periodic", [False, True])
def test_oracle_matches_torch_stft(reflect, periodic):
    for y in _waves([4000, 16000, 1777]):
        ref = torch.stft(torch.from_numpy(y).double(), n_fft=320, hop_length=160, win_length=320,
                         window=torch.hamming_window(320, periodic=periodic, dtype=torch.float64), center=True,
                         pad_mode="reflect" if reflect else "constant", return_complex=True)
        want = torch.log1p(ref.abs()).float()
        got = FO.stft_log_spectrogram(y, reflect=reflect, normalize=False, periodic=periodic)
        assert got.shape == tuple(want.shape) == (161, 1 + len(y) // 160)
        assert rel_err(torch.from_numpy(got), want) < 1e-5
        gotn = torch.from_numpy(FO.stft_log_spectrogram(y, reflect=reflect, normalize=True, periodic=periodic))
        assert rel_err(gotn, (want - want.mean()) / want.std()) < 1e-5          # torch std is unbiased, as data_loader.py:86


def test_oracle_collate_layout():
    inputs, sizes, pct = FO.collate(_waves([3200, 1600]))
    assert inputs.shape == (2, 1, 161, 21) and list(sizes) == [21, 11]
    assert np.all(inputs[1, 0, :, 11:] == 0) and abs(pct[1] - 11 / 21) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [(0, 2e-5), (3, 5e-5), (1, 5e-3)])
@pytest.mark.parametrize("reflect,periodic", [(True, False), (False, False), (True, True)])
def test_gpu_features_match_oracle(precision, tol, reflect, periodic):
    import b200asr
    waves = _waves([16000, 12345, 8000, 700, 160], seed=3)
    want, sizes, pct = FO.collate(waves, reflect=reflect, periodic=periodic)
    got, gpct, gsizes = b200asr.spectrogram_batch([torch.from_numpy(w).cuda() for w in waves], reflect=reflect, precision=precision,
                                                  window_periodic=periodic)
    assert got.shape == want.shape
    assert torch.equal(gsizes.cpu(), torch.from_numpy(sizes)) and rel_err(gpct, torch.from_numpy(pct)) < 1e-6
    assert rel_err(got, torch.from_numpy(want)) < tol
    pad = torch.arange(want.shape[3])[None, :] >= torch.from_numpy(sizes)[:, None].long()
    assert float(got.cpu()[pad[:, None, None, :].expand_as(got)].abs().max()) == 0.0          # zero padded like _collate_fn


@pytest.mark.gpu
def test_gpu_features_feed_the_model_input_contract():
    """The output triple is what Transformer.forward consumes: (B,1,161,T) fp32, lengths in frames."""
    import b200asr
    waves = [torch.from_numpy(w).cuda() for w in _waves([9600, 6400], seed=5)]
    x, pct, sizes = b200asr.spectrogram_batch(waves)
    assert x.shape == (2, 1, 161, 61) and x.dtype == torch.float32 and sizes.tolist() == [61, 41]
    v = x[0, 0]
    assert abs(float(v.mean())) < 1e-4 and abs(float(v.std()) - 1.0) < 1e-4                      # normalised per utterance
