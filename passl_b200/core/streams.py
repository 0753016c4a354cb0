"""Side stream for work that is off the critical path of the backward pass.

Weight-gradient GEMMs only feed the optimizer, while the data-gradient chain (BN backward -> dgrad -> next layer) is what the
next kernel waits for.  The BN kernels are HBM-bound and the wgrad GEMMs are L2 / tensor bound, so running the wgrad of layer l
on a second stream under the BN backward of layer l-1 overlaps two different bottlenecks ("streams").

    with side_stream(x, dy):          # side waits for everything enqueued so far, then runs the block
        K.conv2d_wgrad(...)
    ...
    join()                            # before anything reads the gradients (optimizer, all-reduce, tests)

`record_stream` keeps the caching allocator from recycling the operands while the side stream still reads them.
Disable with PASSL_B200_SIDE_STREAM=0 (serial order on the current stream, used by the parity tests of the flag itself).
"""
import contextlib
import os

import torch

_state = {"stream": None, "pending": False}
ENABLED = os.environ.get("PASSL_B200_SIDE_STREAM", "1") != "0"


def _side():
    if _state["stream"] is None:
        _state["stream"] = torch.cuda.Stream()
    return _state["stream"]


@contextlib.contextmanager
def side_stream(*operands):
    if not ENABLED:
        yield
        return
    side = _side()
    side.wait_stream(torch.cuda.current_stream())
    for t in operands:
        if t is not None:
            t.record_stream(side)
    with torch.cuda.stream(side):
        yield
    _state["pending"] = True


def join():
    """Make the current stream wait for the side stream (no host sync)."""
    if _state["pending"]:
        torch.cuda.current_stream().wait_stream(_state["stream"])
        _state["pending"] = False
