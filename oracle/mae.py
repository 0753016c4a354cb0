"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the MAE pieces of the hot path — numpy float64 restatement of
passl/models/mae.py:156-168 (patchify), :170-182 (unpatchify), :184-212 (random_masking, noise supplied as an input since
Paddle's RNG stream is not reproducible), :268-284 (forward_loss) and passl/models/utils/pos_embed.py:31-82.
Pinned against tests/golden/reference_heads.npz (outputs of the reference source itself)."""
import numpy as np


def patchify(imgs, p=16):
    """imgs (N, 3, H, W) -> (N, L, p*p*3), 'nchpwq->nhwpqc'."""
    n, c, H, W = imgs.shape
    assert H == W and H % p == 0
    h = w = H // p
    x = imgs.reshape(n, 3, h, p, w, p)
    x = np.einsum('nchpwq->nhwpqc', x)
    return x.reshape(n, h * w, p * p * 3)


def unpatchify(x, p=16):
    n, L, _ = x.shape
    h = w = int(L ** .5)
    assert h * w == L
    x = x.reshape(n, h, w, p, p, 3)
    x = np.einsum('nhwpqc->nchpwq', x)
    return x.reshape(n, 3, h * p, h * p)


def random_masking(x, mask_ratio, noise):
    """Returns (x_masked, mask, ids_restore); mask: 0 keep / 1 remove; ids are int64."""
    N, L, D = x.shape
    len_keep = int(L * (1 - mask_ratio))
    ids_shuffle = np.argsort(noise, axis=1, kind="stable")
    ids_restore = np.argsort(ids_shuffle, axis=1, kind="stable")
    ids_keep = ids_shuffle[:, :len_keep]
    x_masked = x[np.arange(N)[:, None], ids_keep]
    mask = np.ones([N, L])
    mask[:, :len_keep] = 0
    mask = mask[np.arange(N)[:, None], ids_restore]
    return x_masked, mask, ids_restore.astype(np.int64)


def forward_loss(imgs, pred, mask, norm_pix_loss=False, p=16):
    target = patchify(imgs.astype(np.float64), p)
    if norm_pix_loss:
        mean = target.mean(axis=-1, keepdims=True)
        var = target.var(axis=-1, keepdims=True, ddof=1)       # paddle Tensor.var is unbiased
        target = (target - mean) / (var + 1.e-6) ** .5
    loss = (pred - target) ** 2
    loss = loss.mean(axis=-1)
    return (loss * mask).sum() / mask.sum()


def forward_loss_grad(imgs, pred, mask, norm_pix_loss=False, p=16):
    """d loss / d pred (pred is the only differentiable input: SURVEY.md App. B)."""
    target = patchify(imgs.astype(np.float64), p)
    if norm_pix_loss:
        mean = target.mean(axis=-1, keepdims=True)
        var = target.var(axis=-1, keepdims=True, ddof=1)
        target = (target - mean) / (var + 1.e-6) ** .5
    return 2.0 * (pred - target) / pred.shape[-1] * mask[..., None] / mask.sum()


def get_1d_sincos_pos_embed_from_grid(embed_dim, pos):
    assert embed_dim % 2 == 0
    omega = np.arange(embed_dim // 2, dtype=np.float32)
    omega /= embed_dim / 2.
    omega = 1. / 10000 ** omega
    pos = pos.reshape(-1)
    out = np.einsum('m,d->md', pos, omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def get_2d_sincos_pos_embed(embed_dim, grid_size, cls_token=False):
    grid_h = np.arange(grid_size, dtype=np.float32)
    grid_w = np.arange(grid_size, dtype=np.float32)
    grid = np.stack(np.meshgrid(grid_w, grid_h), axis=0).reshape([2, 1, grid_size, grid_size])
    emb_h = get_1d_sincos_pos_embed_from_grid(embed_dim // 2, grid[0])
    emb_w = get_1d_sincos_pos_embed_from_grid(embed_dim // 2, grid[1])
    pos_embed = np.concatenate([emb_h, emb_w], axis=1)
    if cls_token:
        pos_embed = np.concatenate([np.zeros([1, embed_dim]), pos_embed], axis=0)
    return pos_embed
