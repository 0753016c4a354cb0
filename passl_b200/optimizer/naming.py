"""Paddle-style automatic parameter names for the modules of this package.

Why: the reference's LARS reads `exclude_from_weight_decay` as a list of SUBSTRINGS of `param.name`
(paddle.fluid.optimizer.LarsMomentumOptimizer, registered at passl_v110/solver/optimizer.py:25), and `param.name` is Paddle's
generated name — `<layer scope>_<n>.w_0` / `.b_0`, the layer scope being the snake-cased class (`conv2d`, `batch_norm2d`,
`batch_norm1d`, `linear`, `layer_norm`) — not the structured `state_dict` key.  configs/moco_byol/moco_byol_r50_IM.yaml:122
(`['batch_norm', '.b_0']`) is written against those names; configs/simclr/simclr_r50_IM.yaml:120 (`["scale","offset",".bias"]`)
matches none of them, so the reference's SimCLR recipe applies LARS scaling and decay to every tensor.  To give the same YAML the
same effect, the optimizers here match the list against names generated the same way.

The counter `<n>` follows the order in which this package registers its modules, not Paddle's construction order; it only
matters for patterns that spell out an index.
"""
import re

_SCOPE = {"ConvBN": "conv2d", "Stem": "conv2d", "Linear": "linear", "LayerNorm": "layer_norm"}


def _snake(name):
    s = re.sub(r"(.)([A-Z][a-z]+)", r"\1_\2", name)
    return re.sub(r"([a-z])([A-Z])", r"\1_\2", s).lower()


def paddle_auto_names(module):
    """Names aligned with `module.parameters()` / ParamStore.params."""
    counters, by_id = {}, {}

    def scope_of(mod, parent):
        cls = type(mod).__name__
        if cls == "BatchNormState":                      # the BatchNorm of a ConvBN / Stem (2-d) or of a BatchNorm1D wrapper
            return "batch_norm1d" if type(parent).__name__ == "BatchNorm1D" else "batch_norm2d"
        return _SCOPE.get(cls, _snake(cls))

    def visit(mod, parent):
        own = list(mod.named_parameters(recurse=False))
        if own:
            scope = scope_of(mod, parent)
            n = counters.get(scope, 0)
            counters[scope] = n + 1
            extra = 0
            for pname, p in own:
                if pname == "bias":
                    suffix = "b_0"
                elif pname == "weight":
                    suffix = "w_0"
                else:                                    # create_parameter() tensors (cls_token, pos_embed, ...): w_0, w_1, ...
                    suffix = "w_%d" % extra
                    extra += 1
                by_id.setdefault(id(p), "%s_%d.%s" % (scope, n, suffix))
        for child in mod.children():
            visit(child, mod)
    visit(module, None)
    return [by_id[id(p)] for p in module.parameters()]
