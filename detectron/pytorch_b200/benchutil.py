"""Host-side helpers of the benchmark harness: algorithmic-bytes accounting,
multi-rank aggregation (one process per GPU, max-over-ranks timing), clock sampling.

No oracle imports here: this is product-side host logic and is covered by CPU tests
(tests/test_benchutil.py, incl. a world_size-2 gloo run of `aggregate`).
"""
import os
import statistics
import subprocess
import threading

import numpy as np


# ------------------------------------------------------------------------------------------------
# algorithmic bytes
# ------------------------------------------------------------------------------------------------
def _f32(x):
    return np.asarray(x, dtype=np.float32)


def _fma32(a, b, c):
    # fp32 fused multiply-add emulated through fp64 (the product of two fp32 is exact in fp64)
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def roi_align_touched_cells(rois, N, H, W, PH, PW, scale, sr):
    """Number of distinct (n, y, x) cells hit by any bilinear tap of the Caffe2-exact RoIAlign
    (identical for every channel).  numpy restatement of the sampling rule with the kernel's
    rounding; used only for the bytes model."""
    rois = _f32(rois)
    scale = np.float32(scale)
    hit = np.zeros((N, H, W), dtype=bool)
    for roi in rois:
        b = int(roi[0])
        if b < 0 or b >= N:
            continue
        sw, sh = roi[1] * scale, roi[2] * scale
        rw = np.maximum(_fma32(roi[3:4], _f32([scale]), -_f32([sw]))[0], np.float32(1))
        rh = np.maximum(_fma32(roi[4:5], _f32([scale]), -_f32([sh]))[0], np.float32(1))
        bh, bw = np.float32(rh / np.float32(PH)), np.float32(rw / np.float32(PW))
        gh = sr if sr > 0 else int(np.ceil(bh))
        gw = sr if sr > 0 else int(np.ceil(bw))

        def axis(start, binsz, P, g, size):
            p = np.repeat(np.arange(P), g).astype(np.float32)
            i = np.tile(np.arange(g), P).astype(np.float32)
            base = _fma32(p, np.full_like(p, binsz), np.full_like(p, start))
            off = (((i + np.float32(.5)) * binsz) / np.float32(g)).astype(np.float32)
            v = (base + off).astype(np.float32)
            valid = ~((v < -1.0) | (v > size))
            v = np.maximum(v, 0)
            low = v.astype(np.int64)
            low = np.minimum(low, size - 1)
            high = np.minimum(low + 1, size - 1)
            return low[valid], high[valid]

        yl, yh = axis(sh, bh, PH, gh, H)
        xl, xh = axis(sw, bw, PW, gw, W)
        ys = np.unique(np.concatenate([yl, yh]))
        xs = np.unique(np.concatenate([xl, xh]))
        if len(ys) and len(xs):
            # every (y-sample, x-sample) pair is taken, so the touched set is a cross product
            hit[b][np.ix_(ys, xs)] = True
    return int(hit.sum())


def roi_align_bytes(shape, R, PH, PW, touched_cells=None):
    """Algorithmic HBM bytes of one RoIAlign forward / backward launch."""
    N, C, H, W = shape
    whole = N * C * H * W * 4
    out = R * C * PH * PW * 4
    rois = R * 5 * 4
    fwd_read = whole if touched_cells is None else min(whole, touched_cells * C * 4)
    return dict(fwd=fwd_read + out + rois, bwd=out + whole + rois)


def nms_bytes(n, kept):
    cb = (n + 63) // 64
    return n * 5 * 4 + 2 * n * cb * 8 + 4 * kept


# ------------------------------------------------------------------------------------------------
# multi-rank aggregation (one process per GPU; no data-path collective: images shard by rank)
# ------------------------------------------------------------------------------------------------
def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    return rank, world, local


def aggregate(elapsed_ms, units, device=None):
    """(max elapsed over ranks, sum of units over ranks).  Uses the default process group if one is
    initialised (NCCL on the GPU box, gloo in the CPU tests), else returns the inputs."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(elapsed_ms), float(units)
    t = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=device)
    u = torch.tensor([float(units)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t.item()), float(u.item())


def shard_images(num_images, rank, world):
    """Contiguous, balanced shard of image indices for this rank (data parallel by image)."""
    base, rem = divmod(num_images, world)
    start = rank * base + min(rank, rem)
    return list(range(start, start + base + (1 if rank < rem else 0)))


# ------------------------------------------------------------------------------------------------
# NUMA placement of pinned host buffers
# ------------------------------------------------------------------------------------------------
def gpu_local_cpus(device_index=0):
    """CPUs on the NUMA node the GPU's PCIe root hangs off (from sysfs), or None if unknown."""
    try:
        import torch
        props = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        with open("/sys/bus/pci/devices/%s/local_cpulist" % bdf) as f:
            text = f.read().strip()
        cpus = set()
        for part in text.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        return cpus or None
    except Exception:  # noqa: BLE001
        return None


class numa_local(object):
    """Context manager: run on the GPU-local CPUs (so that pinned allocations made inside are
    first-touched on the GPU's NUMA node), then restore the previous affinity."""

    def __init__(self, device_index=0):
        self.cpus = gpu_local_cpus(device_index)
        self.prev = None

    def __enter__(self):
        if self.cpus:
            try:
                self.prev = os.sched_getaffinity(0)
                os.sched_setaffinity(0, self.cpus)
            except OSError:
                self.prev = None
        return self

    def __exit__(self, *exc):
        if self.prev is not None:
            try:
                os.sched_setaffinity(0, self.prev)
            except OSError:
                pass
        return False


# ------------------------------------------------------------------------------------------------
# end-to-end pipeline (host buffers in, host buffers out) used by bench.py's `e2e` leg
# ------------------------------------------------------------------------------------------------
class E2EPipeline(object):
    """fn(features, rois) + backward(dY) through the reference-shaped autograd function with HOST buffers.

    Three streams, double-buffered device tensors: the H2D copy of step i+1, the kernels of step i and the D2H copy
    of step i-1 overlap (PCIe is full duplex); every step still moves all its inputs and all its results.
    `h2d` / `compute` / `d2h` switch the legs off for tools/e2e_probe.py (what bounds the step?)."""

    def __init__(self, fn, shape, R, C, P, device, h_feat, h_rois, h_dy, h_out, h_dx, h2d=True, compute=True, d2h=True):
        import torch
        self.torch = torch
        self.fn = fn
        self.h = (h_feat, h_rois, h_dy, h_out, h_dx)
        self.legs = (h2d, compute, d2h)
        self.s_in, self.s_out = torch.cuda.Stream(), torch.cuda.Stream()
        self.d_feat = [torch.zeros(shape, device=device) for _ in range(2)]
        self.d_rois = [h_rois.to(device) for _ in range(2)]
        self.d_dy = [torch.zeros((R, C, P, P), device=device) for _ in range(2)]
        self.d_out = torch.zeros((R, C, P, P), device=device)       # stand-ins when the compute leg is off
        self.d_dx = torch.zeros(shape, device=device)
        self.ev_in = [torch.cuda.Event() for _ in range(2)]          # inputs of slot s have landed
        self.ev_free = [torch.cuda.Event() for _ in range(2)]        # kernels that read slot s are done
        self.ev_out = [torch.cuda.Event() for _ in range(2)]         # results of slot s are on the host side of the copy
        self.keep = [None, None]                                     # results of slot s, alive until their D2H copy is done

    def step(self, i):
        torch = self.torch
        h_feat, h_rois, h_dy, h_out, h_dx = self.h
        h2d, compute, d2h = self.legs
        s = i % 2
        cur = torch.cuda.current_stream()
        if h2d:
            with torch.cuda.stream(self.s_in):
                if i >= 2:
                    self.s_in.wait_event(self.ev_free[s])
                self.d_feat[s].copy_(h_feat, non_blocking=True)
                self.d_rois[s].copy_(h_rois, non_blocking=True)
                self.d_dy[s].copy_(h_dy, non_blocking=True)
                self.ev_in[s].record(self.s_in)
            cur.wait_event(self.ev_in[s])
        if i >= 2 and d2h:
            # The results of step i-2 are released here, after the compute stream has been ordered behind their D2H
            # copy: the caching allocator then hands the same blocks out again every step (no record_stream, whose
            # deferred frees make the allocator grow / cudaFree under this access pattern).
            cur.wait_event(self.ev_out[s])
        self.keep[s] = None
        if compute:
            F = self.d_feat[s].detach().requires_grad_(True)
            out = self.fn(F, self.d_rois[s])
            out.backward(self.d_dy[s])
            grad = F.grad
            out = out.detach()
        else:
            out, grad = self.d_out, self.d_dx
        self.ev_free[s].record(cur)
        if d2h:
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self.ev_free[s])
                if i >= 1:
                    self.s_out.wait_event(self.ev_out[(i - 1) % 2])      # host result buffers are reused every step
                h_out.copy_(out, non_blocking=True)
                h_dx.copy_(grad, non_blocking=True)
                self.ev_out[s].record(self.s_out)
            self.keep[s] = (out, grad)

    def run(self, n):
        for i in range(n):
            self.step(i)
        cur = self.torch.cuda.current_stream()
        cur.wait_stream(self.s_out)            # the timed region ends when the last result is on the host
        cur.wait_stream(self.s_in)
        self.keep = [None, None]


def pcie_bandwidth(device, h_src, h_dst, reps=5):
    """Copy bandwidth of this box's host link with the e2e leg's own pinned buffers: H2D alone, D2H alone and both
    directions at once (GB/s, aggregate for `duplex`).  Context for the e2e number, which is bound by these."""
    import torch
    d_a = torch.empty_like(h_src, device=device)
    d_b = torch.zeros_like(h_dst, device=device)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    out = {}
    for mode in ("h2d", "d2h", "duplex"):
        nbytes = 0
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream()
        e0.record()
        s1.wait_stream(cur); s2.wait_stream(cur)
        for _ in range(reps):
            if mode != "d2h":
                with torch.cuda.stream(s1):
                    d_a.copy_(h_src, non_blocking=True)
                nbytes += h_src.numel() * h_src.element_size()
            if mode != "h2d":
                with torch.cuda.stream(s2):
                    h_dst.copy_(d_b, non_blocking=True)
                nbytes += h_dst.numel() * h_dst.element_size()
        cur.wait_stream(s1); cur.wait_stream(s2)
        e1.record()
        torch.cuda.synchronize()
        out[mode + "_gbs"] = nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9
    return out


# ------------------------------------------------------------------------------------------------
# clocks / throttle reasons during the timed region
# ------------------------------------------------------------------------------------------------
_QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
          "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
          "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")


class ClockSampler(object):
    """Runs `nvidia-smi --query-gpu=... -lms 100` for one GPU while a timed region executes."""

    def __init__(self, gpu_index=0, period_ms=100):
        self.gpu_index = gpu_index
        self.period_ms = period_ms
        self.lines = []
        self.proc = None
        self.thread = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + _QUERY, "--format=csv,noheader,nounits",
                 "-lms", str(self.period_ms)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return self
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()          # exact PID we started
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        if self.thread is not None:
            self.thread.join(timeout=5)
        return False

    def summary(self):
        return summarize_clock_lines(self.lines)


def summarize_clock_lines(lines):
    sm, smax, power = [], [], []
    reasons = set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for line in lines:
        parts = [p.strip() for p in line.split(",")]
        if len(parts) < 9:
            continue
        try:
            sm.append(float(parts[1])); smax.append(float(parts[2])); power.append(float(parts[3]))
        except ValueError:
            continue
        for name, val in zip(names, parts[5:9]):
            if val.lower().startswith("active"):
                reasons.add(name)
    if not sm:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "power_w_max": max(power),
            "reasons": sorted(reasons), "samples": len(sm)}
