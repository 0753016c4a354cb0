"""Golden crop boxes from the reference's OWN sampler, RandomResizedCropAndInterpolationWithTwoPic.get_params
(passl_v110/datasets/preprocess/transforms.py:517-557), called unbound with a stub image under seeded `random` — TEST
INFRASTRUCTURE, build container only:  cd tests/golden && python make_golden_input.py  ->  reference_crop_params.npz.
(The pixel arithmetic of the input stage is pinned against Pillow directly in tests/test_oracle_input_stage_cpu.py.)"""
import importlib
import os
import random
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def main():
    sys.path.insert(0, HERE)
    import make_golden
    import paddle_shim
    make_golden.setup()
    for name, sub in [("passl_v110.datasets", "datasets"), ("passl_v110.datasets.preprocess", "datasets/preprocess")]:
        paddle_shim.fake_package(name, os.path.join(REF, "passl_v110", sub))
    pt = importlib.import_module("paddle.vision.transforms")          # created on demand by the shim's finder

    class _Base:                                              # stands in for every paddle.vision transform class the file registers / extends
        def __init__(self, *a, **k):
            pass
    for n in ("RandomResizedCrop", "ColorJitter", "Transpose", "Normalize", "RandomHorizontalFlip", "Resize", "CenterCrop", "ToTensor",
              "BaseTransform", "RandomCrop", "Compose"):
        setattr(pt, n, type(n, (_Base,), {}))
    tr = importlib.import_module("passl_v110.datasets.preprocess.transforms")
    cls = tr.RandomResizedCropAndInterpolationWithTwoPic
    out = {}
    cases = {"a": (500, 375, (0.08, 1.0), (3. / 4., 4. / 3.)), "b": (320, 480, (0.2, 1.0), (3. / 4., 4. / 3.)),
             "c": (224, 224, (0.1, 1.0), (3. / 4., 4. / 3.)), "d": (640, 64, (0.9, 1.0), (0.9, 1.1)),       # d, e: fallback branches
             "e": (48, 600, (0.95, 1.0), (0.95, 1.05)), "f": (300, 300, (0.75, 1.0), (1.0, 1.0))}
    for tag, (W, H, scale, ratio) in cases.items():
        random.seed(1000 + ord(tag))
        stub = types.SimpleNamespace(size=(W, H))
        boxes = [cls.get_params(None, stub, scale, ratio) for _ in range(64)]
        out["args_" + tag] = np.array([W, H, scale[0], scale[1], ratio[0], ratio[1], 1000 + ord(tag)], dtype=np.float64)
        out["boxes_" + tag] = np.array(boxes, dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "reference_crop_params.npz"), **out)
    print("wrote reference_crop_params.npz", {k: v.shape for k, v in out.items() if k.startswith("boxes")})


if __name__ == "__main__":
    main()
