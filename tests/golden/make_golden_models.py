"""Golden vectors for the necks: run the REFERENCE'S OWN neck classes
(/root/reference/passl_v110/modeling/necks/base_neck.py:43-94,209-237, read in place, never copied) over the torch-backed paddle
shim with seeded weights, and store inputs / weights / outputs in tests/golden/reference_necks.npz.

    python tests/golden/make_golden_models.py          (build container only; the GPU box has no /root/reference)

Paddle layer semantics restated for the shim (paddle 2.4 docs): nn.Linear stores weight [in, out] and computes x @ W + b;
nn.BatchNorm1D in training mode normalises with the biased batch variance, epsilon 1e-5; AdaptiveAvgPool2D((1,1)) = mean over H, W;
fluid.layers.squeeze(x, axes=[]) removes all size-1 dims; fluid.layers.l2_normalize(x, axis, epsilon=1e-12) = x / sqrt(sum x^2 + eps).
The reference initialisers are replaced by explicit seeded weights (stored in the file), so no Paddle RNG is involved."""
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import paddle_shim  # noqa: E402
import make_golden  # noqa: E402


def extend_shim():
    nn = sys.modules["paddle.nn"]
    layers = sys.modules["paddle.fluid.layers"]

    class Linear(nn.Layer):
        def __init__(self, in_features, out_features, **kw):
            super().__init__()
            self.weight = torch.nn.Parameter(torch.zeros(in_features, out_features, dtype=torch.float64))   # paddle: [in, out]
            self.bias = torch.nn.Parameter(torch.zeros(out_features, dtype=torch.float64))

        def forward(self, x):
            return x @ self.weight + self.bias

    class ReLU(nn.Layer):
        def forward(self, x):
            return torch.relu(x)

    class BatchNorm1D(nn.Layer):
        def __init__(self, num_features, momentum=0.9, epsilon=1e-5, **kw):
            super().__init__()
            self.weight = torch.nn.Parameter(torch.ones(num_features, dtype=torch.float64))
            self.bias = torch.nn.Parameter(torch.zeros(num_features, dtype=torch.float64))
            self.epsilon = epsilon

        def forward(self, x):                                   # training mode: batch statistics, biased variance
            mean = x.mean(dim=0)
            var = x.var(dim=0, unbiased=False)
            return (x - mean) / torch.sqrt(var + self.epsilon) * self.weight + self.bias

    class AdaptiveAvgPool2D(nn.Layer):
        def __init__(self, output_size):
            super().__init__()
            assert tuple(output_size) == (1, 1)

        def forward(self, x):
            return x.mean(dim=(2, 3), keepdim=True)

    class LayerNorm(nn.Layer):
        def __init__(self, normalized_shape, epsilon=1e-5, **kw):
            super().__init__()
            d = normalized_shape if isinstance(normalized_shape, int) else normalized_shape[-1]
            self.weight = torch.nn.Parameter(torch.ones(d, dtype=torch.float64))
            self.bias = torch.nn.Parameter(torch.zeros(d, dtype=torch.float64))
            self.epsilon = epsilon

        def forward(self, x):
            mean = x.mean(dim=-1, keepdim=True)
            var = x.var(dim=-1, unbiased=False, keepdim=True)
            return (x - mean) / torch.sqrt(var + self.epsilon) * self.weight + self.bias

    class GELU(nn.Layer):                                      # paddle.nn.GELU(approximate=False): exact erf form
        def forward(self, x):
            return 0.5 * x * (1.0 + torch.erf(x / 2 ** 0.5))

    class Dropout(nn.Layer):
        def __init__(self, p=0.0, **kw):
            super().__init__()
            assert p == 0.0, "the pretrain configs of the hot path use rate 0"

        def forward(self, x):
            return x

    class Identity(nn.Layer):
        def forward(self, x):
            return x

    nn.LayerNorm, nn.GELU, nn.Dropout, nn.Identity = LayerNorm, GELU, Dropout, Identity
    _Lin = Linear

    class LinearB(_Lin):                                       # bias_attr=False -> no bias
        def __init__(self, in_features, out_features, bias_attr=None, **kw):
            super().__init__(in_features, out_features)
            if bias_attr is False:
                self.bias = None

        def forward(self, x):
            y = x @ self.weight
            return y if self.bias is None else y + self.bias
    Linear = LinearB
    _reshape = torch.Tensor.reshape
    torch.Tensor.matmul = lambda self, other: self @ other
    nn.Linear, nn.ReLU, nn.BatchNorm1D, nn.AdaptiveAvgPool2D = Linear, ReLU, BatchNorm1D, AdaptiveAvgPool2D
    for name in ("BatchNorm2D", "GroupNorm", "SyncBatchNorm", "Conv2D"):
        setattr(nn, name, type(name, (nn.Layer,), {}))
    layers.squeeze = lambda x, axes: x.squeeze() if not axes else x.squeeze(*axes)
    # the initialisers are replaced by explicit seeded weights below
    init = types.ModuleType("passl_v110.modules.init")
    for fn in ("init_backbone_weight", "normal_init", "kaiming_init", "constant_", "reset_parameters", "xavier_init",
               "init_backbone_weight_simclr"):
        setattr(init, fn, lambda *a, **k: None)
    sys.modules["passl_v110.modules.init"] = init
    torch.nn.Module.sublayers = lambda self: list(self.modules())[1:]


def seeded(module, rng):
    """Overwrite every parameter with seeded values; returns {name: array} (paddle layout)."""
    out = {}
    for name, p in module.named_parameters():
        if p.dim() == 2:
            v = rng.randn(*p.shape) / np.sqrt(p.shape[0])
        elif name.endswith("weight"):
            v = 1.0 + 0.2 * rng.randn(*p.shape)
        else:
            v = 0.1 * rng.randn(*p.shape)
        with torch.no_grad():
            p.copy_(torch.from_numpy(v))
        out[name] = v
    return out


def main():
    make_golden.setup()
    extend_shim()
    mod = importlib.import_module("passl_v110.modeling.necks.base_neck")
    rng = np.random.RandomState(2024)
    out = {}
    # NonLinearNeckV1 (MoCo v2): avgpool -> fc -> relu -> fc
    neck = mod.NonLinearNeckV1(in_channels=32, hid_channels=48, out_channels=16, with_avg_pool=True)
    for k, v in seeded(neck, rng).items():
        out["v1_" + k] = v
    x = rng.randn(6, 32, 3, 3)
    out["v1_x"], out["v1_y"] = x, neck(torch.from_numpy(x)).detach().numpy()
    # LinearNeck: avgpool -> fc
    neck = mod.LinearNeck(in_channels=32, out_channels=16, with_avg_pool=True)
    for k, v in seeded(neck, rng).items():
        out["lin_" + k] = v
    out["lin_x"], out["lin_y"] = x, neck(torch.from_numpy(x)).detach().numpy()
    # NonLinearNeckfc3 (SimCLR): fc-bn-relu, fc-bn-relu, fc-bn, l2_normalize
    neck = mod.NonLinearNeckfc3(in_channels=32, hid_channels=40, out_channels=24)
    for k, v in seeded(neck, rng).items():
        out["fc3_" + k] = v
    x3 = rng.randn(10, 32, 1, 1)
    out["fc3_x"], out["fc3_y"] = x3, neck(torch.from_numpy(x3)).detach().numpy()
    np.savez_compressed(os.path.join(HERE, "reference_necks.npz"), **out)
    print("wrote reference_necks.npz with", sorted(out))


if __name__ == "__main__":
    main()


def gen_mocov3_loss():
    """MoCoV3Pretrain.contrastive_loss (passl/models/mocov3.py:187-198) called as an unbound method on a stub `self` (T,
    concat_all_gather) — world size 1 in the shim, so labels are arange(N); stored in tests/golden/reference_mocov3.npz."""
    mod = importlib.import_module("passl.models.mocov3")
    out = {}
    rng = np.random.RandomState(77)
    for tag, (N, D, T) in {"a": (12, 32, 0.2), "b": (33, 256, 1.0)}.items():
        q, k = rng.randn(N, D), 0.5 * rng.randn(N, D)
        k[: N // 2] += q[: N // 2]
        stub = types.SimpleNamespace(T=T)
        stub.concat_all_gather = lambda t_: mod.MoCoV3Pretrain.concat_all_gather(stub, t_)
        loss = mod.MoCoV3Pretrain.contrastive_loss(stub, torch.from_numpy(q), torch.from_numpy(k))
        out["q_" + tag], out["k_" + tag], out["T_" + tag], out["loss_" + tag] = q, k, np.float64(T), loss.detach().numpy()
    np.savez_compressed(os.path.join(HERE, "reference_mocov3.npz"), **out)
    print("wrote reference_mocov3.npz", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if k.startswith("loss")})


def gen_vit_block():
    """passl/models/vision_transformer.py:159-206 `Block` (Attention :116-156, Mlp :84-113) with seeded weights:
    tests/golden/reference_vit_block.npz (weights in the paddle layout [in, out])."""
    vt = importlib.import_module("passl.models.vision_transformer")

    class _NoInit(types.ModuleType):                           # initialisers replaced by explicit seeded weights
        def __getattr__(self, n):
            return lambda *a, **k: None
    vt.init = _NoInit("init")
    rng = np.random.RandomState(5)
    out = {}
    for tag, (D, H, N) in {"a": (64, 4, 7), "b": (96, 3, 50)}.items():
        blk = vt.Block(dim=D, num_heads=H, mlp_ratio=4, qkv_bias=True, epsilon=1e-6)
        for k, v in seeded(blk, rng).items():
            out["%s_%s" % (tag, k)] = v
        x = rng.randn(3, N, D)
        out[tag + "_x"], out[tag + "_y"] = x, blk(torch.from_numpy(x)).detach().numpy()
        out[tag + "_heads"] = np.int64(H)
    np.savez_compressed(os.path.join(HERE, "reference_vit_block.npz"), **out)
    print("wrote reference_vit_block.npz", sorted(k for k in out if k.startswith("a_")))


def gen_clip_block():
    """v110 `Block` (passl_v110/modeling/backbones/vision_transformer.py:141-183: QuickGELU MLP, additive attention mask) with the
    causal mask of clip.py:293-295 and without a mask: tests/golden/reference_clip_block.npz."""
    import paddle                                              # the shim
    F_ = sys.modules["paddle.nn.functional"]
    F_.sigmoid = torch.sigmoid
    for pk in ("paddle.nn", "paddle.fluid", "paddle.nn.functional", "paddle.fluid.layers"):
        sys.modules[pk].__path__ = []                            # let `<pkg>.<sub>` imports resolve to placeholders
    lay = paddle_shim._Permissive("paddle.nn.layer")
    lay.__path__ = []
    tr = paddle_shim._Permissive("paddle.nn.layer.transformer")
    tr._convert_attention_mask = lambda mask, dtype: mask        # paddle: converts bool / int masks to additive float; float stays
    sys.modules["paddle.nn.layer"], sys.modules["paddle.nn.layer.transformer"] = lay, tr
    vt = importlib.import_module("passl_v110.modeling.backbones.vision_transformer")
    vt._convert_attention_mask = tr._convert_attention_mask
    rng = np.random.RandomState(11)
    out = {}
    for tag, (D, H, N, causal) in {"causal": (64, 2, 9, True), "plain": (64, 4, 13, False)}.items():
        mask = torch.triu(torch.full((N, N), float("-inf"), dtype=torch.float64), 1) if causal else None
        blk = vt.Block(dim=D, num_heads=H, mlp_ratio=4.0, qkv_bias=True, attn_mask=mask, epsilon=1e-5)
        for k, v in seeded(blk, rng).items():
            out["%s_%s" % (tag, k)] = v
        x = rng.randn(2, N, D)
        out[tag + "_x"], out[tag + "_y"] = x, blk(torch.from_numpy(x)).detach().numpy()
        out[tag + "_heads"] = np.int64(H)
    np.savez_compressed(os.path.join(HERE, "reference_clip_block.npz"), **out)
    print("wrote reference_clip_block.npz", sorted(k for k in out if k.startswith("causal_"))[:6], "...")


def gen_resnet_layer():
    """The reference's vendored ResNet code (passl_v110/modeling/backbones/resnetimagenet.py:93-246): a reduced-width network built
    by the reference's own `ResNet._make_layer` + `BottleneckBlock` (stem conv7x7/2 + BN + ReLU + max-pool, then one stage of 2
    blocks at stride 1 and one stage of 2 blocks at stride 2, widths 8 / 16 so the weights fit in a fixture), train-mode BatchNorm:
    tests/golden/reference_resnet_layers.npz.  Paddle semantics restated in the shim: Conv2D = cross-correlation with zero padding,
    weight [Cout, Cin, R, S]; BatchNorm2D train mode = biased batch variance, epsilon 1e-5; MaxPool2D(3, 2, 1) pads with -inf."""
    nn = sys.modules["paddle.nn"]
    Fn = torch.nn.functional

    class Conv2D(nn.Layer):
        def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias_attr=None, **kw):
            super().__init__()
            assert groups == 1 and dilation == 1 and bias_attr is False
            self.weight = torch.nn.Parameter(torch.zeros(out_channels, in_channels, kernel_size, kernel_size, dtype=torch.float64))
            self.stride, self.padding = stride, padding

        def forward(self, x):
            return Fn.conv2d(x, self.weight, stride=self.stride, padding=self.padding)

    class BatchNorm2D(nn.Layer):
        def __init__(self, num_features, momentum=0.9, epsilon=1e-5, **kw):
            super().__init__()
            self.weight = torch.nn.Parameter(torch.ones(num_features, dtype=torch.float64))
            self.bias = torch.nn.Parameter(torch.zeros(num_features, dtype=torch.float64))
            self.epsilon = epsilon

        def forward(self, x):
            mean = x.mean(dim=(0, 2, 3), keepdim=True)
            var = x.var(dim=(0, 2, 3), unbiased=False, keepdim=True)
            return (x - mean) / torch.sqrt(var + self.epsilon) * self.weight.view(1, -1, 1, 1) + self.bias.view(1, -1, 1, 1)

    class MaxPool2D(nn.Layer):
        def __init__(self, kernel_size, stride, padding):
            super().__init__()
            self.k, self.s, self.p = kernel_size, stride, padding

        def forward(self, x):
            return Fn.max_pool2d(x, self.k, self.s, self.p)
    nn.Conv2D, nn.BatchNorm2D, nn.MaxPool2D = Conv2D, BatchNorm2D, MaxPool2D
    rn = importlib.import_module("passl_v110.modeling.backbones.resnetimagenet")

    class Tiny(rn.ResNet):                                   # same methods, reduced widths: only __init__'s channel table changes
        def __init__(self):
            nn.Layer.__init__(self)
            self.num_classes, self.with_pool, self._norm_layer = 0, False, nn.BatchNorm2D
            self.inplanes, self.dilation = 16, 1
            self.conv1 = nn.Conv2D(3, self.inplanes, kernel_size=7, stride=2, padding=3, bias_attr=False)
            self.bn1 = self._norm_layer(self.inplanes)
            self.relu = nn.ReLU()
            self.maxpool = nn.MaxPool2D(kernel_size=3, stride=2, padding=1)
            self.layer1 = self._make_layer(rn.BottleneckBlock, 8, 2)
            self.layer2 = self._make_layer(rn.BottleneckBlock, 16, 2, stride=2)

        def forward(self, x):
            x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
            return self.layer2(self.layer1(x))
    rng = np.random.RandomState(31)
    net = Tiny()
    out = {"w_" + k: v for k, v in seeded_conv(net, rng).items()}
    x = rng.randn(3, 3, 40, 40)
    out["x"], out["y"] = x, net(torch.from_numpy(x)).detach().numpy()
    np.savez_compressed(os.path.join(HERE, "reference_resnet_layers.npz"), **out)
    print("wrote reference_resnet_layers.npz:", len(out) - 2, "parameters, y", out["y"].shape)


def seeded_conv(module, rng):
    out = {}
    for name, p in module.named_parameters():
        if p.dim() == 4:
            v = rng.randn(*p.shape) / np.sqrt(p.shape[1] * p.shape[2] * p.shape[3])
        elif name.endswith("weight"):
            v = 1.0 + 0.2 * rng.randn(*p.shape)
        else:
            v = 0.1 * rng.randn(*p.shape)
        with torch.no_grad():
            p.copy_(torch.from_numpy(v))
        out[name] = v
    return out


def gen_mae_model():
    """The reference's whole MaskedAutoencoderViT (passl/models/mae.py:37-290: PatchEmbed, pos-embeds, random_masking with the
    noise supplied through paddle.rand, encoder, decoder, forward_loss) at a reduced size, seeded weights, norm_pix_loss on and off:
    tests/golden/reference_mae_model.npz (weights in the paddle layouts: Linear [in, out], Conv2D [E, C, p, p])."""
    import paddle
    nn = sys.modules["paddle.nn"]
    Fn = torch.nn.functional

    class Conv2Db(nn.Layer):                                   # patch embedding: kernel = stride = patch, with bias
        def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias_attr=None, **kw):
            super().__init__()
            k = kernel_size if isinstance(kernel_size, (tuple, list)) else (kernel_size, kernel_size)
            self.weight = torch.nn.Parameter(torch.zeros(out_channels, in_channels, k[0], k[1], dtype=torch.float64))
            self.bias = None if bias_attr is False else torch.nn.Parameter(torch.zeros(out_channels, dtype=torch.float64))
            self.stride, self.padding = stride, padding

        def forward(self, x):
            return Fn.conv2d(x, self.weight, self.bias, stride=self.stride, padding=self.padding)
    nn.Conv2D = Conv2Db
    nn.LayerList = torch.nn.ModuleList
    torch.Tensor._share_buffer_to = lambda self, other: None    # paddle-internal aliasing used only by the initialisers
    nn.Layer.create_parameter = lambda self, shape, **kw: torch.nn.Parameter(torch.zeros(tuple(shape), dtype=torch.float64),
                                                                             requires_grad=False)   # forward-only goldens

    class _NoInit(types.ModuleType):
        def __getattr__(self, n):
            return lambda *a, **k: None
    vt = importlib.import_module("passl.models.vision_transformer")
    vt.init = _NoInit("init")
    mae = importlib.import_module("passl.models.mae")
    mae.init = _NoInit("init")
    rng = np.random.RandomState(19)
    out = {}
    N, IMG, P = 4, 32, 8
    imgs, noise = rng.randn(N, 3, IMG, IMG), rng.rand(N, (IMG // P) ** 2)
    paddle.rand = lambda shape, **kw: torch.from_numpy(noise)
    _orig_no_grad = torch.Tensor.copy_
    for npl in (0, 1):
        net = mae.MaskedAutoencoderViT(img_size=IMG, patch_size=P, in_chans=3, embed_dim=32, depth=2, num_heads=2,
                                       decoder_embed_dim=16, decoder_depth=1, decoder_num_heads=2, mlp_ratio=4.,
                                       norm_pix_loss=bool(npl))
        if npl == 0:
            for name, prm in net.named_parameters():
                if "pos_embed" in name:                        # fixed sin-cos tables from initialize_weights(): keep, and record
                    out["w_" + name] = prm.detach().numpy().copy()
                    continue
                if prm.dim() >= 2:
                    v = rng.randn(*prm.shape) / np.sqrt(np.prod(prm.shape[1:]) if prm.dim() == 4 else prm.shape[0])
                elif name.endswith("weight"):
                    v = 1.0 + 0.2 * rng.randn(*prm.shape)
                else:
                    v = 0.1 * rng.randn(*prm.shape)
                out["w_" + name] = v
        with torch.no_grad():
            for name, prm in net.named_parameters():
                prm.copy_(torch.from_numpy(out["w_" + name]).reshape(prm.shape))
        loss, pred, mask = net(torch.from_numpy(imgs), mask_ratio=0.75)
        out["loss%d" % npl], out["pred%d" % npl], out["mask%d" % npl] = loss.detach().numpy(), pred.detach().numpy(), mask.numpy()
    out["imgs"], out["noise"] = imgs, noise
    np.savez_compressed(os.path.join(HERE, "reference_mae_model.npz"), **out)
    print("wrote reference_mae_model.npz: loss", float(out["loss0"]), float(out["loss1"]), "pred", out["pred0"].shape)


def gen_clip_model():
    """The reference's whole CLIP model (passl_v110/modeling/backbones/clip.py:184-338 with the v110 VisionTransformer /
    Transformer of backbones/vision_transformer.py) + CLIPHead (heads/clip_head.py:27-35) at a reduced size with seeded weights:
    image / text features, both logit matrices, the three losses and the clamped logit_scale — tests/golden/reference_clip_model.npz."""
    import paddle
    nn = sys.modules["paddle.nn"]

    class Embedding(nn.Layer):
        def __init__(self, num_embeddings, embedding_dim, **kw):
            super().__init__()
            self.weight = torch.nn.Parameter(torch.zeros(num_embeddings, embedding_dim, dtype=torch.float64))

        def forward(self, ids):
            return self.weight[ids]
    nn.Embedding = Embedding
    nn.Layer.add_parameter = lambda self, name, prm: None          # create_parameter results are already attributes
    paddle.shape = lambda x: list(x.shape)
    paddle.get_default_dtype = lambda: "float32"
    tens = paddle_shim._Permissive("paddle.tensor")
    tens.triu = torch.triu
    paddle.tensor = tens
    sys.modules["paddle.tensor"] = tens
    _norm = torch.Tensor.norm
    torch.Tensor.norm = lambda self, p=2, axis=None, keepdim=False, **kw: _norm(self, p=p, dim=axis if axis is not None else kw.get("dim"),
                                                                               keepdim=keepdim)
    clip = importlib.import_module("passl_v110.modeling.backbones.clip")
    head_mod = importlib.import_module("passl_v110.modeling.heads.clip_head")
    cfg = dict(embed_dim=16, image_resolution=32, vision_layers=2, vision_width=64, vision_patch_size=8, pre_norm=True, proj=True,
               patch_bias=False, context_length=8, vocab_size=50, transformer_width=32, transformer_heads=2, transformer_layers=2,
               qkv_bias=True)
    net = clip.CLIP(**cfg)
    rng = np.random.RandomState(41)
    out = {}
    with torch.no_grad():
        for name, prm in net.named_parameters():
            if name == "logit_scale":
                v = np.array([np.log(1 / 0.07)])
            elif prm.dim() >= 2:
                fan = np.prod(prm.shape[1:]) if prm.dim() == 4 else prm.shape[-2] if name.endswith("weight") and prm.dim() == 2 else prm.shape[-1]
                v = rng.randn(*prm.shape) / np.sqrt(fan)
            elif name.endswith("weight"):
                v = 1.0 + 0.2 * rng.randn(*prm.shape)
            else:
                v = 0.1 * rng.randn(*prm.shape)
            prm.copy_(torch.from_numpy(np.asarray(v, dtype=np.float64)).reshape(prm.shape))
            out["w_" + name] = np.asarray(v, dtype=np.float64).reshape(tuple(prm.shape))
    n = 6
    img = rng.randn(n, 3, 32, 32)
    text = rng.randint(1, 49, size=(n, 8)).astype(np.int64)
    text[np.arange(n), rng.randint(1, 8, size=n)] = 49              # EOT = highest id (clip.py:306)
    out["img"], out["text"] = img, text
    ti, tt = torch.from_numpy(img), torch.from_numpy(text)
    out["image_features"] = net.encode_image(ti).detach().numpy()
    out["text_features"] = net.encode_text(tt).detach().numpy()
    il, tl = net(ti, tt, is_train=True)
    out["image_logits"], out["text_logits"] = il.detach().numpy(), tl.detach().numpy()
    out["logit_scale_after"] = net.logit_scale.detach().numpy().copy()
    labels = torch.arange(n)
    o = head_mod.CLIPHead()(il, tl, labels, labels)
    for k in ("img_loss", "text_loss", "loss"):
        out[k] = o[k].detach().numpy()
    for k, v in cfg.items():
        out["cfg_" + k] = np.int64(v)
    np.savez_compressed(os.path.join(HERE, "reference_clip_model.npz"), **out)
    print("wrote reference_clip_model.npz: loss", float(out["loss"]), "features", out["image_features"].shape, out["text_features"].shape)


if __name__ == "__main__":
    gen_mocov3_loss()
    gen_vit_block()
    gen_clip_block()
    gen_resnet_layer()
    gen_mae_model()
    gen_clip_model()


def gen_mocov3_pos_embed():
    """MoCoV3ViT.build_2d_sincos_position_embedding (passl/models/mocov3.py:67-91) called as an unbound method on a stub `self`
    (patch_embed sizes, embed_dim, create_parameter) — the fixed position table of MoCo v3, which is NOT the MAE table (meshgrid
    indexing differs); square and non-square grids -> tests/golden/reference_mocov3_pos.npz."""
    import paddle
    paddle.meshgrid = lambda *xs: torch.meshgrid(*xs, indexing="ij")          # paddle.meshgrid is 'ij'-indexed
    paddle.sin, paddle.cos = torch.sin, torch.cos
    mod = importlib.import_module("passl.models.mocov3")
    out = {}
    for tag, (ih, iw, p, dim) in {"a": (112, 112, 8, 64), "b": (32, 32, 8, 768), "c": (24, 40, 8, 32)}.items():
        box = {}

        class _P:
            def set_value(self, v):
                box["v"] = v
        stub = types.SimpleNamespace(patch_embed=types.SimpleNamespace(img_size=(ih, iw), patch_size=(p, p)), embed_dim=dim,
                                     create_parameter=lambda shape: _P())
        mod.MoCoV3ViT.build_2d_sincos_position_embedding(stub)
        out["cfg_" + tag], out["pos_" + tag] = np.array([ih // p, iw // p, dim]), box["v"].numpy()
    np.savez_compressed(os.path.join(HERE, "reference_mocov3_pos.npz"), **out)
    print("wrote reference_mocov3_pos.npz", {k: v.shape for k, v in out.items() if k.startswith("pos")})


if __name__ == "__main__":
    gen_mocov3_pos_embed()
