"""CPU (build container only): the drop-in claim, checked against the reference itself.

north_star: "keep the Python/PyTorch operator surface ... so train_mm_vi_model1.py and translate_mm_vi.py drop it in
unchanged".  The reference is imported UNMODIFIED from /root/reference through oracle/ref_shims.py and, for every symbol
its two entry scripts and its trainer touch on this path (train_mm_vi_model1.py:208-271,420-452; translate_mm_vi.py:100-136;
onmt/TrainerMultimodal.py:166-181,342-346,576-587,625-718; onmt/translate/TranslatorMultimodalVI.py:125-138,199), the
mirrored class must accept the reference's positional / keyword arguments in the reference's order (extra trailing
parameters are allowed only with defaults) and expose the attributes those callers read.  Skipped where /root/reference
does not exist (the GPU box)."""
import inspect
import pickle

import pytest

from oracle import ref_shims

pytestmark = pytest.mark.skipif(not ref_shims.available(), reason="the reference tree is only present in the build container")


@pytest.fixture(scope="module")
def ref():
    return ref_shims.load()


@pytest.fixture(scope="module")
def vm():
    import variational_mmt_b200
    return variational_mmt_b200


def _params(fn):
    return [p for p in inspect.signature(fn).parameters.values() if p.name != "self"]


def _accepts_reference_call(ours, theirs, what):
    po, pt = _params(ours), _params(theirs)
    kinds = (inspect.Parameter.POSITIONAL_ONLY, inspect.Parameter.POSITIONAL_OR_KEYWORD)
    to = [p for p in po if p.kind in kinds]
    tt = [p for p in pt if p.kind in kinds]
    ours_has_kwargs = any(p.kind == inspect.Parameter.VAR_KEYWORD for p in po)
    ours_has_args = any(p.kind == inspect.Parameter.VAR_POSITIONAL for p in po)
    for i, p in enumerate(tt):
        if i < len(to):
            assert to[i].name == p.name, f"{what}: positional #{i} is '{to[i].name}' here, '{p.name}' in the reference"
            if p.default is not inspect.Parameter.empty:
                assert to[i].default is not inspect.Parameter.empty, f"{what}: '{p.name}' is optional in the reference"
        else:
            assert ours_has_args or ours_has_kwargs, f"{what}: reference parameter '{p.name}' is missing"
    for p in to[len(tt):]:
        assert p.default is not inspect.Parameter.empty, f"{what}: extra parameter '{p.name}' has no default"
    if any(p.kind == inspect.Parameter.VAR_KEYWORD for p in pt):
        assert ours_has_kwargs, f"{what}: the reference takes **kwargs"


def test_model_constructor_and_model(ref, vm):
    o = ref.onmt
    _accepts_reference_call(vm.make_vi_model_mmt, o.ModelConstructor.make_vi_model_mmt, "make_vi_model_mmt")
    _accepts_reference_call(vm.NMTVIModel.forward, o.Models.NMTVIModel.forward, "NMTVIModel.forward")
    _accepts_reference_call(vm.NMTVIModel.__init__, o.Models.NMTVIModel.__init__, "NMTVIModel.__init__")
    _accepts_reference_call(vm.RNNEncoder.__init__, o.Models.RNNEncoder.__init__, "RNNEncoder.__init__")
    _accepts_reference_call(vm.RNNEncoder.forward, o.Models.RNNEncoder.forward, "RNNEncoder.forward")


def test_decoder_and_state(ref, vm):
    o = ref.onmt
    from onmt.VI_Model1 import StdRNNVIModel1Decoder as RefDec
    _accepts_reference_call(vm.StdRNNVIModel1Decoder.__init__, RefDec.__init__, "StdRNNVIModel1Decoder.__init__")
    _accepts_reference_call(vm.StdRNNVIModel1Decoder.forward, RefDec.forward, "decoder.forward")
    _accepts_reference_call(vm.StdRNNVIModel1Decoder.init_decoder_state, RefDec.init_decoder_state, "init_decoder_state")
    _accepts_reference_call(vm.StdRNNVIModel1Decoder._run_forward_pass, RefDec._run_forward_pass, "_run_forward_pass")
    R = o.Models.RNNDecoderState
    for m in ("__init__", "update_state", "repeat_beam_size_times", "beam_update", "detach"):
        _accepts_reference_call(getattr(vm.RNNDecoderState, m), getattr(R, m), "RNNDecoderState." + m)
    assert isinstance(getattr(vm.RNNDecoderState, "_all"), property)


def test_attention_and_inference_networks(ref, vm):
    o = ref.onmt
    G = o.modules.GlobalAttention
    _accepts_reference_call(vm.GlobalAttention.__init__, G.__init__, "GlobalAttention.__init__")
    _accepts_reference_call(vm.GlobalAttention.forward, G.forward, "GlobalAttention.forward")
    import onmt.modules.NormalVariationalEncoder as N
    for cls in ("LocationLayer", "ScaleLayer", "GlobalInferenceNetwork", "GlobalFullInferenceNetwork",
                "ImageGlobalInferenceNetwork"):
        _accepts_reference_call(getattr(vm, cls).__init__, getattr(N, cls).__init__, cls + ".__init__")
        _accepts_reference_call(getattr(vm, cls).forward, getattr(N, cls).forward, cls + ".forward")
    _accepts_reference_call(vm.GlobalInferenceNetwork.encode_seq, N.GlobalInferenceNetwork.encode_seq, "encode_seq")
    import onmt.modules.Dists as D
    for m in ("params", "mean", "sample"):
        assert callable(getattr(vm.Normal, m)) and hasattr(D.Normal, m)


def test_loss_compute_and_statistics(ref, vm):
    o = ref.onmt
    L = o.VILoss.NMTVIModel1LossCompute
    _accepts_reference_call(vm.NMTVIModel1LossCompute.__init__, L.__init__, "NMTVIModel1LossCompute.__init__")
    for m in ("sharded_compute_loss", "monolithic_compute_loss", "_compute_loss", "_make_shard_state"):
        _accepts_reference_call(getattr(vm.NMTVIModel1LossCompute, m), getattr(L, m), "NMTVIModel1LossCompute." + m)
    S = o.VIStatistics                              # onmt/TrainerMultimodal.py:32, re-exported as onmt.VIStatistics
    _accepts_reference_call(vm.VIStatistics.__init__, S.__init__, "VIStatistics.__init__")
    for m in ("update", "accuracy", "ppl", "elapsed_time", "output"):
        assert callable(getattr(vm.VIStatistics, m, None)), "VIStatistics." + m
        _accepts_reference_call(getattr(vm.VIStatistics, m), getattr(S, m), "VIStatistics." + m)
    # the attributes TrainerMultimodal and EarlyStop read from a statistics object (VILoss.py:484-496 keys)
    ours, theirs = vm.VIStatistics("vi-model1"), S("vi-model1")
    for a in vars(theirs):
        assert hasattr(ours, a), "VIStatistics." + a


def test_optim_contract_and_checkpoint_pickle(ref, vm):
    o = ref.onmt
    R = o.Optim
    _accepts_reference_call(vm.Optim.__init__, R.__init__, "Optim.__init__")
    for m in ("set_parameters", "step", "update_learning_rate", "_set_rate"):
        _accepts_reference_call(getattr(vm.Optim, m), getattr(R, m), "Optim." + m)
    ours = vm.Optim("adam", 0.002, 5, lr_decay=0.5, start_decay_at=8)
    theirs = R("adam", 0.002, 5, lr_decay=0.5, start_decay_at=8)
    for a in vars(theirs):                      # lr, original_lr, max_grad_norm, method, _step, betas, ...
        assert hasattr(ours, a), "Optim." + a
    # TrainerMultimodal.drop_checkpoint pickles the whole object ('optim': self.optim, TrainerMultimodal.py:576-587) and
    # train_mm_vi_model1.build_optim reads optim.optimizer.state_dict() back (:433-437)
    ours._step, ours.lr = 17, 0.001
    back = pickle.loads(pickle.dumps(ours))
    assert (back._step, back.lr, back.method, back.max_grad_norm) == (17, 0.001, "adam", 5)
    back.optimizer.load_state_dict(back.optimizer.state_dict())
    assert back.optimizer.param_groups[0]["lr"] == 0.001


def test_translator_contract(ref, vm):
    o = ref.onmt
    T = o.translate.TranslatorMultimodalVI
    _accepts_reference_call(vm.TranslatorMultimodalVI.__init__, T.__init__, "TranslatorMultimodalVI.__init__")
    _accepts_reference_call(vm.TranslatorMultimodalVI.translate_batch, T.translate_batch, "translate_batch")
    _accepts_reference_call(vm.GNMTGlobalScorer.__init__, o.translate.GNMTGlobalScorer.__init__, "GNMTGlobalScorer.__init__")


def test_state_dict_keys_match_the_reference_model(ref, vm):
    """Checkpoints move both ways: the mirrored model has exactly the reference's parameter names and shapes."""
    from oracle import synth
    from variational_mmt_b200 import synthetic
    for cfg in (synth.TINY, synth.TINY_FIXED):
        params = synth.make_params(cfg, 3435, 0.1)
        rmodel, _ = ref.build_model(cfg, params)
        opt = synthetic.make_opt(emb=cfg.emb, hidden=cfg.hidden, z_dim=cfg.z_dim, layers=cfg.layers,
                                 conditional=cfg.conditional, dropout=0.0)
        ours = vm.make_vi_model_mmt(opt, synthetic.make_fields(cfg.v_src, cfg.v_tgt), gpu=False)
        a = {k: tuple(v.shape) for k, v in rmodel.state_dict().items()}
        b = {k: tuple(v.shape) for k, v in ours.state_dict().items()}
        assert a == b
