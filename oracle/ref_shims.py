"""Import the UNMODIFIED reference (/root/reference) under torch 2.x with runtime shims.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Only usable in the build container (the reference is
not present on the GPU box); used by oracle/make_golden.py to pin the restatement in
oracle/vi_model1_ref.py / oracle/beam_ref.py.  No reference file is edited or copied: every shim is
a monkeypatch applied from outside (list and rationale: SURVEY.md section 8c).
"""
import argparse
import contextlib
import io
import os
import sys
import types
import warnings

REF = os.environ.get("VMMT_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "onmt"))


_loaded = {}


def load():
    """Returns a namespace with torch, onmt, opts and helpers; idempotent."""
    if _loaded:
        return _loaded["ns"]
    import torch

    # 1. absent third-party modules imported at module scope by onmt/io and the train script
    tt, ttd, ttv = (types.ModuleType(n) for n in ("torchtext", "torchtext.data", "torchtext.vocab"))

    class _Stub:
        def __init__(self, *a, **k):
            pass
    ttd.Dataset = ttd.Iterator = ttd.Field = ttd.Example = _Stub
    ttv.Vocab = type("Vocab", (), {})
    tt.data, tt.vocab = ttd, ttv
    sys.modules.update({"torchtext": tt, "torchtext.data": ttd, "torchtext.vocab": ttv,
                        "tables": types.ModuleType("tables")})
    # 2. onmt/Utils.py:5-9 asserts on a relative METEOR jar path at import time
    real_isfile = os.path.isfile
    os.path.isfile = lambda q: True if str(q).endswith("meteor-1.5.jar") else real_isfile(q)
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir(REF)
    warnings.filterwarnings("ignore")
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            import onmt
            import onmt.ModelConstructor
            import onmt.Utils
            import onmt.VILoss
            import onmt.translate
            import opts
    finally:
        os.chdir(cwd)
        os.path.isfile = real_isfile

    # 3. `1 - mask` on a bool mask (GlobalAttention.py:176)
    class BoolMask(torch.Tensor):
        def __rsub__(self, other):
            return ~self.as_subclass(torch.Tensor)
    orig_sm = onmt.Utils.sequence_mask
    sys.modules["onmt.modules.GlobalAttention"].sequence_mask = \
        lambda lengths, max_len=None: orig_sm(lengths, max_len).as_subclass(BoolMask)

    # 4. torch.stack(Tensor) (Models.py:1151-1154)
    orig_stack = torch.stack

    def stack(tensors, *a, **k):
        if isinstance(tensors, torch.Tensor):
            tensors = list(tensors.unbind(0))
        return orig_stack(tensors, *a, **k)
    torch.stack = stack

    # 5. Normal.std (Dists.py:19)
    import torch.distributions as td
    if not hasattr(td.Normal, "std"):
        td.Normal.std = property(lambda self: self.scale)

    # 7. legacy aliasing semantics of compute_cosine (hazard H3, VILoss.py:22-56): in-place
    #    normalisation of the shared storage, autograd history not rebased.
    def legacy_compute_cosine(pred, obs):
        dim = pred.dim() - 1
        with torch.no_grad():
            pred.data.div_(pred.data.pow(2).sum(dim).sqrt().unsqueeze(dim))
            obs.data.div_(obs.data.pow(2).sum(dim).sqrt().unsqueeze(dim))
        return torch.nn.functional.cosine_similarity(pred.detach(), obs.detach(), dim=dim)
    onmt.VILoss.compute_cosine = legacy_compute_cosine

    # 8. integer `/` must floor (Beam.py:104)
    orig_truediv = torch.Tensor.__truediv__

    def legacy_truediv(self, other):
        if not self.is_floating_point() and (isinstance(other, int) or (
                isinstance(other, torch.Tensor) and not other.is_floating_point())):
            return torch.div(self, other, rounding_mode="floor")
        return orig_truediv(self, other)
    torch.Tensor.__truediv__ = legacy_truediv

    ns = types.SimpleNamespace(torch=torch, onmt=onmt, opts=opts)

    class FakeVocab:
        def __init__(self, n, specials):
            self.itos = specials + ["w%d" % i for i in range(n - len(specials))]
            self.stoi = {w: i for i, w in enumerate(self.itos)}

        def __len__(self):
            return len(self.itos)

    class FakeField:
        def __init__(self, v):
            self.vocab = v

    def make_fields(vs, vt):
        return {"src": FakeField(FakeVocab(vs, ["<unk>", "<blank>"])),
                "tgt": FakeField(FakeVocab(vt, ["<unk>", "<blank>", "<s>", "</s>"]))}

    def make_opt(cfg, extra=()):
        prs = argparse.ArgumentParser()
        opts.model_opts(prs); opts.train_opts(prs); opts.train_mm_vi_model1_opts(prs)
        args = ["-data", "x", "-path_to_train_img_feats", "resnet50.hdf5",
                "-path_to_valid_img_feats", "v.hdf5", "--multimodal_model_type", "vi-model1",
                "--z_latent_dim", str(cfg.z_dim), "--use_global_image_features",
                "-batch_size", "40", "-optim", "adam", "-learning_rate", "0.002",
                "-rnn_type", "LSTM", "-rnn_size", str(cfg.hidden),
                "-src_word_vec_size", str(cfg.emb), "-tgt_word_vec_size", str(cfg.emb),
                "-layers", str(cfg.layers), "-dropout", str(cfg.dropout), "-dropout_imgs", "0.5"]
        if cfg.conditional:
            args.append("--conditional")
        if getattr(cfg, "brnn", False):
            args += ["-encoder_type", "brnn"]
        opt = prs.parse_args(args + list(extra))
        opt.brnn = (opt.encoder_type == "brnn")
        return opt

    def build_model(cfg, params):
        """make_vi_model_mmt (ModelConstructor.py:328-620) + load our deterministic weights."""
        fields = make_fields(cfg.v_src, cfg.v_tgt)
        with contextlib.redirect_stdout(io.StringIO()):
            model = onmt.ModelConstructor.make_vi_model_mmt(make_opt(cfg), fields, False, None)
        sd = model.state_dict()
        with torch.no_grad():
            for k, t in sd.items():
                src = k.replace("encoder_tgt.embeddings", "decoder.embeddings")
                t.copy_(torch.as_tensor(params[src]))
        return model, fields

    def make_loss(model, fields):
        """NMTVIModel1LossCompute with shim 6 (criterion result of shape [1], VILoss.py:243,478,485)."""
        loss = onmt.VILoss.NMTVIModel1LossCompute(model.generator, fields["tgt"].vocab)
        crit = loss.criterion

        class Crit1(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.c = crit

            def forward(self, a, b):
                return self.c(a, b).view(1)
        loss.criterion = Crit1()
        return loss

    @contextlib.contextmanager
    def inject_noise(eps):
        """z = mu + sigma*eps instead of torch.normal's global RNG (Dists.py:21-26)."""
        from onmt.modules.Dists import Normal
        orig = Normal.sample
        Normal.sample = lambda self: (self.normal.mean + self.normal.scale * eps).detach()
        try:
            yield
        finally:
            Normal.sample = orig

    ns.make_fields, ns.make_opt, ns.build_model, ns.make_loss, ns.inject_noise = \
        make_fields, make_opt, build_model, make_loss, inject_noise
    _loaded["ns"] = ns
    return ns
