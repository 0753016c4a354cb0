"""Parameter-name and layout mapping between this package and the reference's Paddle checkpoints:
`to_paddle_state(model)` / `load_paddle_state(model, state)` for MoCo, SimCLR, MAE, CLIP and MoCo v3 (and their trunks), and
`save_pdparams` / `load_pdparams` for the file container.  ResNet path first; the ViT families follow further down.

A `.pdparams` file written by `paddle.save(layer.state_dict(), path)` is a pickle of {structured_name: numpy.ndarray} (plus the
`StructuredToParameterName@@` bookkeeping entry), so it is read and written here with `pickle` alone.  Mapping rules, derived from
the reference classes (passl_v110/modeling/backbones/resnetimagenet.py:93-246, necks/base_neck.py, architectures/moco.py:58-80):

    ours                                   reference (Paddle)                         layout
    stem.weight [64, 152]                  conv1.weight [64, 3, 7, 7]                 (r, s, c)-flattened + zero pad  <->  NCHW kernel
    stem.bn.{weight,bias,_mean,_variance}  bn1.{weight,bias,_mean,_variance}
    blocks.i.conv{k}.weight [O, R, S, I]   layer{L}.{j}.conv{k}.weight [O, I, R, S]
    blocks.i.conv{k}.bn.*                  layer{L}.{j}.bn{k}.*
    blocks.i.downsample.weight / .bn.*     layer{L}.{j}.downsample.0.weight / downsample.1.*
    neck fc{1,2}.weight [out, in]          mlp.{0,2}.weight [in, out]                 (NonLinearNeckV1; LinearNeck: fc; fc3: mlp.0,3,6 + BN mlp.1,4,7)
    queue [K, dim]                         queue [dim, K]
    queue_ptr int64 [1]                    queue_ptr int64 [1]

The key sets and shapes this produces are checked against the reference classes themselves (tests/test_checkpoint_cpu.py builds
ResNet-50 + NonLinearNeckV1 and MoCoV3Pretrain over the paddle shim; MAE / CLIP names come from the reference models' own
named_parameters() recorded in tests/golden).  Files written by a real Paddle run could not be tested here.
"""
import pickle

import numpy as np
import torch

LAYERS = (3, 4, 6, 3)


def _block_names(layers=LAYERS):
    out, bi = {}, 0
    for li, n in enumerate(layers, start=1):
        for j in range(n):
            out[bi] = "layer%d.%d" % (li, j)
            bi += 1
    return out


def _bn(dst, src, sd, out):
    for k in ("weight", "bias", "_mean", "_variance"):
        if src + "." + k in sd:
            out[dst + "." + k] = sd[src + "." + k].detach().cpu().numpy()


def resnet_to_paddle(backbone, prefix=""):
    """passl_b200 ResNet -> {paddle name: ndarray}."""
    sd = backbone.state_dict()
    out = {}
    w = sd["stem.weight"].detach().cpu()[:, :147].reshape(64, 7, 7, 3).permute(0, 3, 1, 2).contiguous()
    out[prefix + "conv1.weight"] = w.numpy()
    _bn(prefix + "bn1", "stem.bn", sd, out)
    names = _block_names(tuple(getattr(backbone, "_layers", backbone.LAYER_CFG[50])))
    for bi, ref in names.items():
        for k in (1, 2, 3):
            key = "blocks.%d.conv%d" % (bi, k)
            out["%s%s.conv%d.weight" % (prefix, ref, k)] = sd[key + ".weight"].detach().cpu().permute(0, 3, 1, 2).contiguous().numpy()
            _bn("%s%s.bn%d" % (prefix, ref, k), key + ".bn", sd, out)
        key = "blocks.%d.downsample" % bi
        if key + ".weight" in sd:
            out["%s%s.downsample.0.weight" % (prefix, ref)] = sd[key + ".weight"].detach().cpu().permute(0, 3, 1, 2).contiguous().numpy()
            _bn("%s%s.downsample.1" % (prefix, ref), key + ".bn", sd, out)
    return out


def resnet_from_paddle(backbone, state, prefix=""):
    sd = backbone.state_dict()
    new = {}
    w = torch.zeros_like(sd["stem.weight"])
    w[:, :147] = torch.as_tensor(state[prefix + "conv1.weight"]).permute(0, 2, 3, 1).reshape(64, 147)
    new["stem.weight"] = w
    for k in ("weight", "bias", "_mean", "_variance"):
        new["stem.bn." + k] = torch.as_tensor(state[prefix + "bn1." + k])
    for bi, ref in _block_names(tuple(getattr(backbone, "_layers", backbone.LAYER_CFG[50]))).items():
        for k in (1, 2, 3):
            key = "blocks.%d.conv%d" % (bi, k)
            new[key + ".weight"] = torch.as_tensor(state["%s%s.conv%d.weight" % (prefix, ref, k)]).permute(0, 2, 3, 1).contiguous()
            for s in ("weight", "bias", "_mean", "_variance"):
                new[key + ".bn." + s] = torch.as_tensor(state["%s%s.bn%d.%s" % (prefix, ref, k, s)])
        key = "blocks.%d.downsample" % bi
        if key + ".weight" in sd:
            new[key + ".weight"] = torch.as_tensor(state["%s%s.downsample.0.weight" % (prefix, ref)]).permute(0, 2, 3, 1).contiguous()
            for s in ("weight", "bias", "_mean", "_variance"):
                new[key + ".bn." + s] = torch.as_tensor(state["%s%s.downsample.1.%s" % (prefix, ref, s)])
    missing = set(sd) - set(new)
    assert not missing, "unmapped parameters: %s" % sorted(missing)[:5]
    backbone.load_state_dict({k: v.to(sd[k].dtype) for k, v in new.items()})


_NECK_MAPS = {
    "NonLinearNeckV1": {"fc1": "mlp.0", "fc2": "mlp.2"},
    "LinearNeck": {"fc": "fc"},
    "NonLinearNeckfc3": {"fc1": "mlp.0", "bn1.bn": "mlp.1", "fc2": "mlp.3", "bn2.bn": "mlp.4", "fc3": "mlp.6", "bn3.bn": "mlp.7"},
}


def neck_to_paddle(neck, prefix=""):
    sd, out = neck.state_dict(), {}
    for ours, ref in _NECK_MAPS[type(neck).__name__].items():
        for k, v in sd.items():
            if k.startswith(ours + "."):
                a = v.detach().cpu()
                if a.dim() == 2:
                    a = a.t().contiguous()                     # [out, in] -> paddle [in, out]
                out[prefix + ref + k[len(ours):]] = a.numpy()
    return out


def neck_from_paddle(neck, state, prefix=""):
    sd, new = neck.state_dict(), {}
    for ours, ref in _NECK_MAPS[type(neck).__name__].items():
        for k, v in sd.items():
            if k.startswith(ours + "."):
                a = torch.as_tensor(state[prefix + ref + k[len(ours):]])
                new[k] = (a.t().contiguous() if a.dim() == 2 else a).to(v.dtype)
    neck.load_state_dict(new)


def moco_to_paddle(model):
    """MoCo (architectures/moco.py) -> the reference's state_dict names: encoder_{q,k}.0.* backbone, encoder_{q,k}.1.* neck, queue, queue_ptr."""
    model.flush_queue()
    out = {}
    for enc in ("encoder_q", "encoder_k"):
        seq = getattr(model, enc)
        out.update(resnet_to_paddle(seq[0], enc + ".0."))
        out.update(neck_to_paddle(seq[1], enc + ".1."))
    out["queue"] = model.queue.detach().cpu().t().contiguous().numpy()           # [K, dim] -> [dim, K]
    out["queue_ptr"] = model.queue_ptr.detach().cpu().numpy()
    return out


def moco_from_paddle(model, state):
    for enc in ("encoder_q", "encoder_k"):
        seq = getattr(model, enc)
        resnet_from_paddle(seq[0], state, enc + ".0.")
        neck_from_paddle(seq[1], state, enc + ".1.")
    with torch.no_grad():
        model.queue.copy_(torch.as_tensor(state["queue"]).t())
        model.queue_ptr.copy_(torch.as_tensor(state["queue_ptr"]).to(torch.int64).reshape(1))
    model._pending_keys = None


def save_pdparams(state, path):
    """Same container as paddle.save(state_dict): a protocol-2 pickle of {name: ndarray} with the name-table entry."""
    obj = {k: np.asarray(v) for k, v in state.items()}
    obj["StructuredToParameterName@@"] = {k: k for k in state}
    with open(path, "wb") as f:
        pickle.dump(obj, f, protocol=2)


def load_pdparams(path):
    with open(path, "rb") as f:
        obj = pickle.load(f, encoding="latin1")
    obj.pop("StructuredToParameterName@@", None)
    return {k: (load_nested(v) if isinstance(v, dict) else np.asarray(v)) for k, v in obj.items()}


def load_nested(obj):
    """v110 checkpoints wrap the weights as {'state_dict': {...}, 'epoch': n, ...} (hooks/checkpoint_hook.py:22-49)."""
    obj = dict(obj)
    obj.pop("StructuredToParameterName@@", None)
    return {k: (load_nested(v) if isinstance(v, dict) else np.asarray(v)) for k, v in obj.items()}


# ---------------------------------------------------------------------------------------------------------------------------------
# ViT families (MAE, MoCo v3, CLIP).  Rules, derived from passl/models/vision_transformer.py:91-200 (Attention.qkv/proj, Mlp.fc1/fc2),
# passl/models/mae.py:37-120, passl/models/mocov3.py:110-170 and passl_v110/modeling/backbones/clip.py:184-250:
#
#     ours                                     reference (Paddle)                             layout
#     blocks.i.{qkv,proj}.*                    blocks.i.attn.{qkv,proj}.*                     Linear [out, in] <-> [in, out]
#     blocks.i.fc{1,2}.*                       blocks.i.mlp.fc{1,2}.*                         Linear [out, in] <-> [in, out]
#     patch_embed.proj.weight [E, p*p*C]       patch_embed.proj.weight [E, C, p, p]           (p, q, c)-flattened <-> NCHW kernel
#     CLIP visual.proj.weight [out, width]     visual.proj [width, out]
#     CLIP text.{token_embedding, positional_embedding, blocks, ln_final, text_projection.weight}
#                                              {token_embedding.weight, positional_embedding, transformer.blocks, ln_final,
#                                               text_projection [width, out]}
#     MoCo v3 {base_encoder.vit.*, base_encoder.head.fcs.l / bns.l.bn}
#                                              {base_encoder.*, base_encoder.head.(3l) / (3l+1)}  (nn.Sequential indices)
#     MoCo v3 momentum_encoder.{vit,head}.*    momentum_encoder.model.0.{.,head.}*            (CosineEMA wraps Sequential(encoder, predictor))
# ---------------------------------------------------------------------------------------------------------------------------------
import re

_BLK_ATTN = re.compile(r"(blocks\.\d+)\.(qkv|proj)\.")
_BLK_MLP = re.compile(r"(blocks\.\d+)\.(fc[12])\.")
_REF_ATTN = re.compile(r"(blocks\.\d+)\.attn\.(qkv|proj)\.")
_REF_MLP = re.compile(r"(blocks\.\d+)\.mlp\.(fc[12])\.")


def _vit_name_to_ref(name):
    return _BLK_MLP.sub(r"\1.mlp.\2.", _BLK_ATTN.sub(r"\1.attn.\2.", name))


def _vit_name_from_ref(name):
    return _REF_MLP.sub(r"\1.\2.", _REF_ATTN.sub(r"\1.\2.", name))


def _vit_tensor_to_ref(name, t, in_chans=3):
    a = t.detach().cpu()
    if name.endswith("patch_embed.proj.weight"):
        p = int(round((a.shape[1] // in_chans) ** 0.5))
        return a.reshape(a.shape[0], p, p, in_chans).permute(0, 3, 1, 2).contiguous()
    if a.dim() == 2 and name.endswith(".weight"):
        return a.t().contiguous()
    return a


def _vit_tensor_from_ref(name, v, in_chans=3):
    a = torch.as_tensor(np.asarray(v))
    if name.endswith("patch_embed.proj.weight"):
        return a.permute(0, 2, 3, 1).reshape(a.shape[0], -1).contiguous()
    if a.dim() == 2 and name.endswith(".weight"):
        return a.t().contiguous()
    return a


def vit_to_paddle(module, prefix="", in_chans=3):
    """VisionTransformer / MaskedAutoencoderViT / CLIPVisionTransformer trunk -> {paddle name: ndarray}."""
    return {prefix + _vit_name_to_ref(k): _vit_tensor_to_ref(k, v, in_chans).numpy() for k, v in module.state_dict().items()}


def vit_from_paddle(module, state, prefix="", in_chans=3, strict=True):
    sd, new = module.state_dict(), {}
    for k, cur in sd.items():
        ref = prefix + _vit_name_to_ref(k)
        if ref not in state:
            assert not strict, "checkpoint has no entry %s" % ref
            continue
        t = _vit_tensor_from_ref(k, state[ref], in_chans)
        assert tuple(t.shape) == tuple(cur.shape), (k, tuple(t.shape), tuple(cur.shape))
        new[k] = t.to(cur.dtype)
    module.load_state_dict(new, strict=strict)


mae_to_paddle, mae_from_paddle = vit_to_paddle, vit_from_paddle


_CLIP_TOP = {                                  # ours -> (reference name, transpose)
    "logit_scale": ("logit_scale", False),
    "visual.proj.weight": ("visual.proj", True),
    "text.token_embedding": ("token_embedding.weight", False),
    "text.positional_embedding": ("positional_embedding", False),
    "text.text_projection.weight": ("text_projection", True),
}


def _clip_ref_name(k):
    if k in _CLIP_TOP:
        return _CLIP_TOP[k]
    if k.startswith("text.blocks."):
        return "transformer." + _vit_name_to_ref(k[len("text."):]), None
    if k.startswith("text.ln_final."):
        return k[len("text."):], None
    return _vit_name_to_ref(k), None


def clip_to_paddle(model, in_chans=3):
    """CLIP (models/clip.py) -> the reference CLIP's state_dict names (backbones/clip.py:184-250)."""
    out = {}
    for k, v in model.state_dict().items():
        ref, tr = _clip_ref_name(k)
        if tr is None:
            out[ref] = _vit_tensor_to_ref(k, v, in_chans).numpy()
        else:
            a = v.detach().cpu()
            out[ref] = (a.t().contiguous() if tr else a).numpy()
    return out


def clip_from_paddle(model, state, in_chans=3):
    sd, new = model.state_dict(), {}
    for k, cur in sd.items():
        ref, tr = _clip_ref_name(k)
        if tr is None:
            t = _vit_tensor_from_ref(k, state[ref], in_chans)
        else:
            t = torch.as_tensor(np.asarray(state[ref]))
            t = t.t().contiguous() if tr else t
        t = t.reshape(cur.shape) if t.numel() == cur.numel() and t.dim() != cur.dim() else t
        assert tuple(t.shape) == tuple(cur.shape), (k, tuple(t.shape), tuple(cur.shape))
        new[k] = t.to(cur.dtype)
    model.load_state_dict(new)


def _mlpbn_to_paddle(mlp, prefix, out):
    """MLPBN (fcs.l, bns.l.bn) -> nn.Sequential indices of _build_mlp (mocov3.py:136-158): Linear 3l, BatchNorm1D 3l+1 (ReLU 3l+2).
    Paddle keeps weight / bias entries (ones / zeros, stop_gradient) for the affine-free last BatchNorm."""
    sd = mlp.state_dict()
    for l in range(len(mlp.fcs)):
        out["%s%d.weight" % (prefix, 3 * l)] = sd["fcs.%d.weight" % l].detach().cpu().t().contiguous().numpy()
        if l < len(mlp.bns):
            c = sd["bns.%d.bn._mean" % l].shape[0]
            for s in ("weight", "bias", "_mean", "_variance"):
                key = "bns.%d.bn.%s" % (l, s)
                if key in sd:
                    out["%s%d.%s" % (prefix, 3 * l + 1, s)] = sd[key].detach().cpu().numpy()
                else:
                    out["%s%d.%s" % (prefix, 3 * l + 1, s)] = (np.ones if s == "weight" else np.zeros)(c, dtype=np.float32)


def _mlpbn_from_paddle(mlp, state, prefix):
    sd, new = mlp.state_dict(), {}
    for k, cur in sd.items():
        kind, l, rest = k.split(".", 2)
        ref = "%s%d.%s" % (prefix, 3 * int(l) + (0 if kind == "fcs" else 1), rest[len("bn."):] if kind == "bns" else rest)
        t = torch.as_tensor(np.asarray(state[ref]))
        new[k] = (t.t().contiguous() if kind == "fcs" else t).to(cur.dtype)
    mlp.load_state_dict(new)


def mocov3_to_paddle(model, in_chans=3):
    """MoCoV3Pretrain -> reference names: base_encoder.* (ViT with head = projector Sequential), predictor.*, and the EMA copy under
    momentum_encoder.model.0.* (averaged_model.py:36 keeps `self.model = deepcopy(Sequential(base_encoder, predictor))`).  The
    reference's EMA also averages a predictor copy (model.1.*) and sends the keys through it; in literal mode
    (reference_ema_quirk) that copy exists here and is exported, otherwise the live predictor stands in for it."""
    out = {}
    for ours, ref in (("base_encoder", "base_encoder."), ("momentum_encoder", "momentum_encoder.model.0.")):
        enc = getattr(model, ours)
        out.update(vit_to_paddle(enc.vit, ref, in_chans))
        _mlpbn_to_paddle(enc.head, ref + "head.", out)
    _mlpbn_to_paddle(model.predictor, "predictor.", out)
    _mlpbn_to_paddle(getattr(model, "momentum_predictor", None) or model.predictor, "momentum_encoder.model.1.", out)
    out["momentum_encoder.steps"] = np.asarray(model.steps, dtype=np.int64)
    return out


def mocov3_from_paddle(model, state, in_chans=3):
    for ours, ref in (("base_encoder", "base_encoder."), ("momentum_encoder", "momentum_encoder.model.0.")):
        enc = getattr(model, ours)
        vit_from_paddle(enc.vit, state, ref, in_chans)
        _mlpbn_from_paddle(enc.head, state, ref + "head.")
    _mlpbn_from_paddle(model.predictor, state, "predictor.")
    if getattr(model, "momentum_predictor", None) is not None:
        _mlpbn_from_paddle(model.momentum_predictor, state, "momentum_encoder.model.1.")
    model.steps = int(np.asarray(state.get("momentum_encoder.steps", 0)))


def simclr_to_paddle(model):
    """SimCLR (architectures/simclr.py:43-47): `encoder` = Sequential(backbone, neck) and `backbone` aliases encoder[0]; both
    prefixes are written (a recursive state_dict lists the aliased sub-layer twice)."""
    out = resnet_to_paddle(model.encoder[0], "encoder.0.")
    out.update(neck_to_paddle(model.encoder[1], "encoder.1."))
    out.update(resnet_to_paddle(model.encoder[0], "backbone."))
    return out


def simclr_from_paddle(model, state):
    pre = "encoder.0." if "encoder.0.conv1.weight" in state else "backbone."
    resnet_from_paddle(model.encoder[0], state, pre)
    neck_from_paddle(model.encoder[1], state, "encoder.1.")


def _prefixed(fn, prefix):
    return lambda model, *a: {prefix + k: v for k, v in fn(model, *a).items()}


def _strip(state, prefix):
    return {k[len(prefix):]: v for k, v in state.items() if k.startswith(prefix)}


def _dispatch(model):
    """model class name -> (export, import) pair; wrappers recurse with the reference's attribute name as prefix."""
    name = type(model).__name__
    table = {
        "MoCo": (moco_to_paddle, moco_from_paddle),
        "SimCLR": (simclr_to_paddle, simclr_from_paddle),
        "ResNet": (resnet_to_paddle, resnet_from_paddle),
        "MaskedAutoencoderViT": (mae_to_paddle, mae_from_paddle),
        "VisionTransformer": (vit_to_paddle, vit_from_paddle),
        "MoCoV3ViT": (vit_to_paddle, vit_from_paddle),
        "CLIPVisionTransformer": (vit_to_paddle, vit_from_paddle),
        "CLIP": (clip_to_paddle, clip_from_paddle),
        "MoCoV3Pretrain": (mocov3_to_paddle, mocov3_from_paddle),
    }
    if name in table:
        return table[name]
    if name == "CLIPWrapper":                  # architectures/CLIPWrapper.py:39: self.model = build_backbone(architecture)
        return (lambda m: {"model." + k: v for k, v in clip_to_paddle(m.model).items()},
                lambda m, st: clip_from_paddle(m.model, _strip(st, "model.")))
    raise NotImplementedError("no Paddle checkpoint mapping for %s (supported: %s, CLIPWrapper)" % (name, ", ".join(sorted(table))))


def to_paddle_state(model):
    """{reference structured name: ndarray in the reference's layout} for any of the model families of this package."""
    return _dispatch(model)[0](model)


def load_paddle_state(model, state):
    """Inverse of to_paddle_state: fills `model` from a reference state dict (e.g. load_pdparams(path))."""
    _dispatch(model)[1](model, state)


# ---------------------------------------------------------------------------------------------------------------------------------
# v110 training checkpoints (`epoch_N.pd`, hooks/checkpoint_hook.py:22-49 + engine/trainer.py:419-431): a plain pickle of
# {'epoch': n, 'state_dict': {name: ndarray}, 'optimizer': {...}, 'lr_scheduler': {'last_epoch': .., 'last_lr': ..}}.
# The optimizer entry holds Paddle's accumulator tensors under Paddle-internal names; it is neither written nor read here
# (momentum / Adam moments restart from zero after such a resume — the torch-format checkpoint of engine/trainer.py keeps them).
# ---------------------------------------------------------------------------------------------------------------------------------
def is_paddle_pickle(path):
    """True for the reference's pickle containers, False for this package's torch.save archives (zip: 'PK' magic)."""
    with open(path, "rb") as f:
        return f.read(2) != b"PK"


def save_v110_checkpoint(path, model, epoch, lr_scheduler=None):
    obj = {"epoch": int(epoch), "state_dict": {k: np.asarray(v) for k, v in to_paddle_state(model).items()}}
    if lr_scheduler is not None:
        obj["lr_scheduler"] = dict(lr_scheduler.state_dict())
    with open(path, "wb") as f:
        pickle.dump(obj, f, protocol=2)


def load_v110_checkpoint(path):
    """-> dict with 'state_dict' (always), and 'epoch' / 'lr_scheduler' when the file has them; a bare weights file
    (.pdparams) comes back as {'state_dict': weights}, like the reference's Trainer.load (engine/trainer.py:433-437)."""
    obj = load_pdparams(path)
    if isinstance(obj.get("state_dict"), dict):
        out = {"state_dict": obj["state_dict"]}
        if obj.get("epoch") is not None:
            out["epoch"] = int(np.asarray(obj["epoch"]))
        if isinstance(obj.get("lr_scheduler"), dict):
            out["lr_scheduler"] = {k: np.asarray(v).item() for k, v in obj["lr_scheduler"].items() if np.asarray(v).ndim == 0}
        return out
    return {"state_dict": obj}
