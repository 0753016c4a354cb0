"""Registry entries for the CLIP path: `CLIPWrapper` (MODELS), `CLIP` (BACKBONES), `CLIPHead` (HEADS) — the names
configs/clip/vit-b-32.yaml uses (architectures/CLIPWrapper.py:26, backbones/clip.py:183, heads/clip_head.py:21)."""
from ...models.clip import CLIP, CLIPHead, CLIPWrapper
from ..registry import BACKBONES, HEADS, MODELS

MODELS.register(CLIPWrapper)
BACKBONES.register(CLIP)
HEADS.register(CLIPHead)
