"""Parameter-name and layout mapping between this package and the reference's Paddle checkpoints for the MoCo /
SimCLR ResNet path: `to_paddle_state(model)` / `load_paddle_state(model, state)` and `save_pdparams` / `load_pdparams`.

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

The key set and shapes this produces for ResNet-50 + NonLinearNeckV1 are checked against the reference classes themselves
(tests/test_checkpoint_cpu.py builds them over the paddle shim).  Files written by a real Paddle run could not be tested here.
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
    names = _block_names(tuple(backbone.LAYER_CFG[50]) if not hasattr(backbone, "_layers") else backbone._layers)
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
    for bi, ref in _block_names().items():
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
    return {k: np.asarray(v) for k, v in obj.items()}
