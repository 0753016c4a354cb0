"""`-p / --profiler_options` of the train entry (tools/train.py:30, passl/utils/profiler.py:26-110): the same option string
("batch_range=[50, 60]; profile_path=...; exit_on_finished=true") selects a window of training steps; here the window is marked for
the CUDA tools of this build — `cudaProfilerStart/Stop` (what `ncu --profile-from-start off` and nsys capture ranges key on) plus an
NVTX range per step — instead of Paddle's operator profiler.  `state`, `sorted_key`, `tracer_option` are accepted and have nothing to
select here; `profile_path` is where a one-line summary of the window (steps, wall time) is written."""
import sys
import time


class ProfilerOptions:
    DEFAULTS = dict(batch_range=[10, 20], state="All", sorted_key="total", tracer_option="Default", profile_path="/tmp/profile",
                    exit_on_finished=True)

    def __init__(self, options_str):
        assert isinstance(options_str, str)
        self._options = {k: (list(v) if isinstance(v, list) else v) for k, v in self.DEFAULTS.items()}
        for item in options_str.replace(" ", "").split(";"):
            if not item:
                continue
            key, _, value = item.partition("=")
            if key == "batch_range":
                bounds = [int(v) for v in value.strip("[]").split(",") if v]
                if len(bounds) >= 2 and 0 <= bounds[0] < bounds[1]:            # an invalid range keeps the default, like the reference
                    self._options[key] = bounds[:2]
            elif key == "exit_on_finished":
                self._options[key] = value.lower() in ("yes", "true", "t", "1")
            elif key in ("state", "sorted_key", "tracer_option", "profile_path"):
                self._options[key] = value

    def __getitem__(self, name):
        if self._options.get(name, None) is None:
            raise ValueError("ProfilerOptions does not have an option named %s." % name)
        return self._options[name]


class StepProfiler:
    """One `step()` call per training iteration (the reference's add_profiler_step): opens the window at batch_range[0], closes it at
    batch_range[1] and exits there when exit_on_finished is set."""

    def __init__(self, options_str=None, backend=None):
        self.options = ProfilerOptions(options_str) if options_str else None
        self.step_id, self.active, self._t0 = 0, False, None
        self._backend = backend                                   # tests pass a recorder; default = torch.cuda

    def _cuda(self):
        if self._backend is not None:
            return self._backend
        import torch
        return torch.cuda

    def step(self):
        if self.options is None:
            return
        lo, hi = self.options["batch_range"]
        cuda = self._cuda()
        if self.active:
            cuda.nvtx.range_pop()
        if self.step_id == lo:
            cuda.synchronize()
            cuda.profiler.start()
            self.active, self._t0 = True, time.time()
        elif self.step_id == hi and self.active:
            cuda.synchronize()
            cuda.profiler.stop()
            self.active = False
            with open(self.options["profile_path"], "w") as f:
                f.write("profiled steps [%d, %d): %.3f ms per step (wall)\n" % (lo, hi, (time.time() - self._t0) * 1e3 / (hi - lo)))
            if self.options["exit_on_finished"]:
                sys.exit(0)
        if self.active:
            cuda.nvtx.range_push("step_%d" % self.step_id)
        self.step_id += 1
