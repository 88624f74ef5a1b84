"""The drop-in swap, executed: the reference's own `HiFiGAN` wrapper (reference models/hifigan.py:618-724)
is built here from /root/reference, its `.generator` is replaced by `HiFiGANGenerator.from_reference(...)`,
and everything that does not need arithmetic is compared with the untouched wrapper -- geometry, the 156- and
232-key state_dict (keys, order, shapes, values), what the wrapper's own state_dict / eval / train / children
see, and the complete stdout of `HiFiGAN.forward` with debug_shapes (the wrapper's two lines around the
generator's six: reference tests/test_hifigan_integration.py:184-211).

Runs on the CPU box only (the reference tree does not travel to the GPU box); the device call is replaced by a
zero waveform of the right shape, since the print / shape contract is what is under test here.  The arithmetic
behind the same module is the business of tests/test_parity_gpu.py."""
import contextlib
import io
import os
import sys

import pytest
import torch

import tts_sambert_hifigan_b200 as pkg

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")),
                                reason="reference tree not present (GPU box)")


@pytest.fixture(scope="module")
def ref_hifigan_cls():
    sys.path.insert(0, REF)
    try:
        from models.hifigan import HiFiGAN
    finally:
        sys.path.remove(REF)
    return HiFiGAN


def _build(cls, **kw):
    with contextlib.redirect_stdout(io.StringIO()):
        torch.manual_seed(1234)
        return cls(**kw).eval()


@pytest.mark.parametrize("geometry", [
    {},
    dict(upsample_rates=[5, 5, 4, 2], upsample_kernel_sizes=[10, 10, 8, 4], upsample_initial_channel=256,
         resblock_kernel_sizes=[3, 5], resblock_dilation_sizes=[[1, 2], [2, 6, 3]]),
])
def test_swap_under_reference_wrapper(ref_hifigan_cls, geometry):
    hifigan = _build(ref_hifigan_cls, **geometry)
    old = hifigan.generator
    new = pkg.HiFiGANGenerator.from_reference(old, mode="tf32")
    # geometry and attributes other code reads (reference models/hifigan.py:171-174)
    assert (new.n_mels, new.num_kernels, new.num_upsamples, new.debug_shapes) == \
           (old.n_mels, old.num_kernels, old.num_upsamples, old.debug_shapes)
    g = new._geometry
    assert g["upsample_rates"] == [m.stride[0] for m in old.ups]
    assert g["upsample_kernel_sizes"] == [m.kernel_size[0] for m in old.ups]
    assert g["upsample_initial_channel"] == old.conv_pre.out_channels
    assert g["resblock_dilation_sizes"] == [[c.dilation[0] for c in rb.convs1] for rb in old.mrfs[0].resblocks]
    # plain schema: same keys in the same order, same shapes, same values
    sd_old, sd_new = old.state_dict(), new.state_dict()
    assert list(sd_old) == list(sd_new)
    for k in sd_old:
        assert sd_old[k].shape == sd_new[k].shape and torch.equal(sd_old[k], sd_new[k]), k
    if not geometry:
        assert len(sd_new) == 156
    # the swap itself (reference :681-689 builds it, :719 calls it)
    hifigan.generator = new
    assert hifigan.generator is new and new in list(hifigan.children())
    full = hifigan.state_dict()
    assert [k for k in full if k.startswith("generator.")] == ["generator." + k for k in sd_old]
    hifigan.train(); hifigan.eval()
    assert new.training is False
    # weight-normed schema: the reference applies it on its generator, ours must load it and show the same keys
    with contextlib.redirect_stdout(io.StringIO()), pytest.warns() if False else contextlib.nullcontext():
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            old.apply_weight_norm()
    sd_wn = old.state_dict()
    new2 = pkg.HiFiGANGenerator.from_reference(old)
    assert list(new2.state_dict()) == list(sd_wn)
    if not geometry:
        assert len(sd_wn) == 232
    for k in sd_wn:
        assert torch.equal(new2.state_dict()[k], sd_wn[k]), k
    new2.remove_weight_norm()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        old.remove_weight_norm()
    for k, v in old.state_dict().items():
        assert torch.allclose(new2.state_dict()[k], v, rtol=1e-6, atol=1e-8), k


def test_wrapper_print_contract_with_swapped_generator(ref_hifigan_cls, monkeypatch):
    """stdout of HiFiGAN.forward(mel) with debug_shapes=True must be the same text whichever generator sits in
    the wrapper.  The device call is stubbed with a zero waveform of the contract shape."""
    mel = torch.zeros(2, 80, 10)

    def stdout_of(model):
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf), torch.no_grad():
            wav = model(mel)
        return buf.getvalue().splitlines(), wav

    ref = _build(ref_hifigan_cls, debug_shapes=True)
    want, wav_ref = stdout_of(ref)
    assert any(l.startswith("[HiFiGAN]") for l in want) and any(l.startswith("[HiFiGANGenerator]") for l in want)

    new = pkg.HiFiGANGenerator.from_reference(ref.generator)
    assert new.debug_shapes is True
    monkeypatch.setattr(pkg.HiFiGANGenerator, "_dispatch",
                        lambda self, mel, shapes, mode, stages, frames_last:
                        torch.zeros((mel.shape[0], 1, shapes[-1][2]), dtype=torch.float32))
    ref.generator = new
    got, wav_new = stdout_of(ref)
    assert got == want
    assert wav_new.shape == wav_ref.shape and wav_new.dtype == wav_ref.dtype == torch.float32


def test_decoder_and_length_regulator_swap_under_reference_acoustic_model():
    """the reference SAMBERTAcousticModel (models/acoustic_model.py:171-178, :258) with its
    O(T^2) decoder and its length regulator replaced by the B200 mirrors keeps its own state_dict key for key."""
    sys.path.insert(0, REF)
    try:
        from models.acoustic_model import SAMBERTAcousticModel
    finally:
        sys.path.remove(REF)
    model = _build(SAMBERTAcousticModel)
    before = {k: v.clone() for k, v in model.state_dict().items()}
    old = model.ar_decoder
    model.ar_decoder = pkg.PNCAARDecoder.from_reference(old, verbose=False)
    model.variance_adaptor.length_regulator = pkg.LengthRegulator()
    after = model.state_dict()
    assert list(after) == list(before)
    for k, v in before.items():
        assert torch.equal(after[k], v), k
    assert (model.ar_decoder.d_model, model.ar_decoder.n_mels, model.ar_decoder.n_layers, model.ar_decoder.n_heads,
            model.ar_decoder.chunk_size) == (old.d_model, old.n_mels, old.n_layers, old.n_heads, old.chunk_size)
