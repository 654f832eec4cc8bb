"""Drop-in proof inside the reference's OWN framework code (needs /root/reference: runs in the build container, skipped
on the GPU box): `blocks.sp_layers` is replaced by openasr_b200's module before `frameworks.Speech_Models` is
imported, then Model.create_model -> package -> restore (Speech_Models.py:66-103) run unmodified.  On a machine with
both a GPU and the reference the forward pass through the reference's CTC model is exercised as well."""
import importlib
import importlib.util
import os
import sys

import pytest
import torch

REF = "/root/reference/src"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not mounted")

SP_CONF = {"feature_type": "fbank", "sample_rate": 16000, "num_mel_bins": 40, "use_energy": False,
           "spec_aug": {"freq_mask_num": 2, "freq_mask_width": 27, "time_mask_num": 2, "time_mask_width": 40}}


@pytest.fixture()
def ref_framework(monkeypatch):
    """Import the reference's frameworks package with OUR module registered as blocks.sp_layers."""
    monkeypatch.syspath_prepend(REF)
    for name in [m for m in sys.modules if m == "blocks" or m.startswith("blocks.") or m.startswith("frameworks")]:
        monkeypatch.delitem(sys.modules, name, raising=False)
    import blocks  # the reference's package (encoders, conv_layers, ...)
    ours = importlib.import_module("openasr_b200.blocks.sp_layers")
    monkeypatch.setitem(sys.modules, "blocks.sp_layers", ours)
    monkeypatch.setattr(blocks, "sp_layers", ours, raising=False)
    # third-party packages the reference's utils import at module level but this path never calls (soundfile,
    # editdistance, ...) are stubbed from OUTSIDE; the reference files themselves are untouched
    import types
    for name in ("soundfile", "editdistance", "ctcdecode", "kaldi_io", "tensorboardX"):
        if importlib.util.find_spec(name) is None:
            monkeypatch.setitem(sys.modules, name, types.ModuleType(name))
    sm = None
    for _ in range(12):
        try:
            sm = importlib.import_module("frameworks.Speech_Models")
            break
        except ModuleNotFoundError as e:
            if not e.name or e.name.split(".")[0] in ("blocks", "frameworks", "utils", "loss", "third_party", "dataload"):
                raise
            monkeypatch.setitem(sys.modules, e.name, types.ModuleType(e.name))
            for name in [m for m in sys.modules if m.startswith("frameworks") or m in ("utils", "loss")]:
                monkeypatch.delitem(sys.modules, name, raising=False)
    if sm is None:
        pytest.skip("reference framework not importable here")
    yield sm


def _en_conf(input_dim=40):
    return {"type": "transformer", "input_dim": input_dim, "d_model": 64, "nhead": 2, "dim_feedforward": 128,
            "num_layers": 1, "dropout_rate": 0.1, "activation": "relu", "sub": {"type": "ConvV2", "layer_num": 2}}


def _de_conf():
    return {"type": "transformer", "d_model": 64, "nhead": 2, "num_layers": 1, "encoder_dim": 64, "dim_feedforward": 128,
            "vocab_size": 30, "dropout_rate": 0.1, "activation": "relu"}


def _create(sm, sp_conf, input_dim=40):
    """What Conv_Transformer.create_model does (Speech_Models.py:205-216), statement by statement.  At this revision
    the reference's model constructors are mid-refactor (Conv_Transformer.__init__ calls Framework.__init__ without
    its arguments, Conv_CTC.__init__ skips it), so the model object is the reference's Framework base with the
    reference's own package / restore bound to it -- every line that touches SPLayer is the reference's."""
    import types
    from blocks.sp_layers import SPLayer          # the reference's import statement -> OUR module
    from blocks.encoders import TransformerEncoder
    from blocks.decoders import TransformerDecoder
    import frameworks
    try:
        model = frameworks.Framework(SPLayer(sp_conf), TransformerEncoder(_en_conf(input_dim)), TransformerDecoder(_de_conf()))
    except (KeyError, TypeError) as e:
        pytest.skip("encoder / decoder config of this reference revision differs: %r" % (e,))
    model.package = types.MethodType(sm.Conv_Transformer.package, model)
    model.restore = types.MethodType(sm.Conv_Transformer.restore, model)
    return model


def test_create_package_restore_with_our_splayer(ref_framework):
    sm = ref_framework
    import openasr_b200
    model = _create(sm, dict(SP_CONF))
    assert isinstance(model.splayer, openasr_b200.SPLayer)          # OUR class was picked up by the reference code
    pkg = model.package()
    assert pkg["splayer_config"] == SP_CONF and len(pkg["splayer_state"]) == 0
    model2 = _create(sm, dict(SP_CONF))
    model2.restore(pkg)                                              # key-by-key config comparison + state dicts
    model3 = _create(sm, dict(SP_CONF, num_mel_bins=80), input_dim=80)
    with pytest.raises(ValueError):
        model3.restore(pkg)                                          # "splayer_config mismatch."
    # extension keys: a checkpoint written with them restores; one written WITHOUT them (by the reference's own
    # SPLayer) is rejected by the reference's restore(), which indexes pkg["splayer_config"][key] for every key of
    # the live config -- so old checkpoints are loaded into a module built from THEIR config (no extension keys)
    ext = dict(SP_CONF, cmvn="utterance", dither=1.0)
    model4 = _create(sm, ext)
    with pytest.raises(KeyError):
        model4.restore(pkg)
    _create(sm, ext).restore(model4.package())


@pytest.mark.gpu
def test_forward_through_reference_model(ref_framework):
    sm = ref_framework
    from oracle import frontend_oracle as fo
    model = _create(sm, dict(SP_CONF, dither=0.0)).cuda().eval()
    w, l = fo.synth_batch(3, 8000, 20000, 16000, seed=1)
    with torch.no_grad():
        enc, lens = model.get_encoded(w.cuda(), l.cuda())
    assert torch.isfinite(enc).all() and enc.shape[0] == 3 and enc.shape[-1] == 64
