"""Host-side plumbing shared by the drop-in model classes: the flat fp32 parameter arena the
nn.Parameters are views of, the binding of a caller-constructed torch.optim.Adam to the arena
(so that `optimizer.state` holds what the reference would hold), the device log buffer and the
data-parallel gradient exchange.  All arithmetic happens in libcdgvae_sm100.so.
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from . import dist as _dist

ALIGN = 32  # floats (128 B): keeps every weight 16-B aligned for float4 / TMA access


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _f32c(t, device):
    """Bring a batch tensor to the device as contiguous fp32 (the reference does `.cuda()`, train.py:163-165)."""
    if t.device != device or t.dtype != torch.float32:
        t = t.to(device=device, dtype=torch.float32, non_blocking=True)
    return t.contiguous()


def flow_apply(scm, flow_num, inverse_loop, params, offsets, x, direction, log_determinant=False):
    """`cdg_flow_apply` on a [batch, node] matrix: every node's flow (direction 0) or its inverse (1) in one launch.
    `params`: fp32 device tensor holding the flow parameters, `offsets[j]`: float offset of node j's block in it."""
    dev = _lib.require_cuda(params.device)
    x = _f32c(x, dev)
    if x.dim() != 2 or x.shape[1] != len(offsets):
        raise ValueError(f"flow input must be [batch, {len(offsets)}], got {tuple(x.shape)}")
    out = torch.empty_like(x)
    logdet = torch.empty_like(x) if log_determinant else None
    off = (C.c_int64 * len(offsets))(*[int(o) for o in offsets])
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cdg_flow_apply(_lib.SCM[scm], int(flow_num), int(inverse_loop), len(offsets), _ptr(params), off,
                                             _ptr(x), x.stride(0), _ptr(out), out.stride(0), _ptr(logdet),
                                             0 if logdet is None else logdet.stride(0), x.shape[0], int(direction),
                                             C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return out, logdet


class ArenaModule(nn.Module):
    """nn.Module whose parameters live in one flat arena (params / grads / exp_avg / exp_avg_sq)."""

    _arena = None

    # -- arena ---------------------------------------------------------------------------------
    def _group_of(self, name):
        # parameters of one flow node are packed contiguously (the kernels index them as a block)
        if name.startswith("flows."):
            return ".".join(name.split(".")[:2])
        return name

    def _arena_named_parameters(self):
        """Parameters that live in the (trainable) arena; the CelebA model keeps its frozen encoder elsewhere."""
        return list(self.named_parameters())

    def _build_arena(self):
        named = self._arena_named_parameters()
        if not named:
            return
        device = named[0][1].device
        off, offsets, prev = 0, {}, None
        for name, p in named:
            g = self._group_of(name)
            if g != prev:
                off = (off + ALIGN - 1) // ALIGN * ALIGN
            offsets[name] = off
            off += p.numel()
            prev = g
        total = (off + ALIGN - 1) // ALIGN * ALIGN
        arena = torch.zeros(total, dtype=torch.float32, device=device)
        grads = torch.zeros_like(arena)
        with torch.no_grad():
            for name, p in named:
                o, n = offsets[name], p.numel()
                arena[o:o + n].copy_(p.data.reshape(-1).to(torch.float32))
                p.data = arena[o:o + n].view(p.shape)
                p.grad = None
        self._arena, self._grads = arena, grads
        self._exp_avg = self._exp_avg_sq = None
        self._offsets, self._n_params = offsets, total
        self._bound_opt = None
        self._destroy_plan()
        self._workspace = None
        self._logs = None
        self.__dict__["_graphs"] = {}
        self._dev_step = None

    def _apply(self, fn, *a, **k):
        super()._apply(fn, *a, **k)
        self._build_arena()
        return self

    def _destroy_plan(self):
        pass

    @property
    def arena_device(self):
        return self._arena.device

    def _lin(self, prefix):
        w, b = prefix + ".weight", prefix + ".bias"
        shape = dict(self._arena_named_parameters())[w].shape
        return _lib.Linear(self._offsets[w], self._offsets[b], shape[1], shape[0])

    def _flow_offsets(self, node):
        first = "p" if self.config["scm"] == "linear" else "w.0"
        return [self._offsets[f"flows.{i}.{first}"] for i in range(node)]

    def _grad_views(self, live_names=None):
        """Expose gradients the way autograd would: p.grad is a view of the gradient arena."""
        for name, p in self._arena_named_parameters():
            if live_names is not None and name not in live_names:
                p.grad = None
                continue
            o, n = self._offsets[name], p.numel()
            p.grad = self._grads[o:o + n].view(p.shape)

    # -- workspace / logs ----------------------------------------------------------------------
    def _get_workspace(self, nbytes):
        ws = self._workspace
        if ws is None or ws.numel() < nbytes:
            self._workspace = ws = torch.zeros(int(nbytes), dtype=torch.uint8, device=self.arena_device)
            self.drop_graphs()                      # captured graphs hold the old workspace address
        return ws

    def _log_rows(self, n_rows, width):
        lg = self._logs
        if lg is None or lg.shape[0] < n_rows or lg.shape[1] != width:
            new = torch.zeros(max(n_rows, 256 if lg is None else 2 * lg.shape[0]), width, device=self.arena_device)
            if lg is not None and lg.shape[1] == width:
                new[: lg.shape[0]].copy_(lg)
            self._logs = lg = new
        return lg

    # -- optimizer -----------------------------------------------------------------------------
    def live_param_names(self):
        """Parameters that receive a gradient in the reference (all, except covtype's unused decoder)."""
        return [n for n, _ in self._arena_named_parameters()]

    def adam_segments(self):
        """(offset, length) ranges of the arena the optimizer touches."""
        segs = []
        shapes = {n: p.numel() for n, p in self._arena_named_parameters()}
        for n in self.live_param_names():
            segs.append((self._offsets[n], shapes[n]))
        return self._merge(segs)

    @staticmethod
    def _merge(segs):
        out = []
        for o, ln in sorted(s for s in segs if s[1] > 0):
            if out and 0 <= o - (out[-1][0] + out[-1][1]) < ALIGN:     # only alignment padding between
                out[-1] = (out[-1][0], o + ln - out[-1][0])
            else:
                out.append((o, ln))
        return out

    def bind_optimizer(self, opt):
        """Make `opt.state[p]['exp_avg'|'exp_avg_sq']` views of the arena-shaped moment buffers and
        remember the hyper-parameters' source.  torch.optim.Adam only (the reference's optimizer)."""
        if self._bound_opt is opt and self._still_bound(opt):
            return
        if self._bound_opt is opt:
            # optimizer.load_state_dict() (or anything else) replaced the state tensors after binding: copy the new
            # moments / step counts into the arenas below and forget graphs captured with the old device step
            self.drop_graphs()
        if not isinstance(opt, torch.optim.Adam) or isinstance(opt, torch.optim.AdamW):
            raise TypeError("the reference trains with torch.optim.Adam; got %s" % type(opt).__name__)
        mine = {id(p): n for n, p in self._arena_named_parameters()}
        groups = [g for g in opt.param_groups if any(id(p) in mine for p in g["params"])]
        if not groups:
            raise ValueError("optimizer does not hold this model's parameters")
        for g in groups[1:]:
            for key in ("lr", "betas", "eps", "weight_decay", "amsgrad", "maximize"):
                if g.get(key) != groups[0].get(key):
                    raise ValueError("parameter groups with different hyper-parameters are not supported")
        if groups[0].get("amsgrad") or groups[0].get("maximize"):
            raise ValueError("amsgrad / maximize are not used by the reference and are not supported")
        self._opt_group = groups[0]
        if self._exp_avg is None:
            self._exp_avg = torch.zeros_like(self._arena)
            self._exp_avg_sq = torch.zeros_like(self._arena)
        step = 0
        self._step_tensors = []
        live = set(self.live_param_names())
        for name, p in self._arena_named_parameters():
            if name not in live:
                continue
            st = opt.state[p]
            o, n = self._offsets[name], p.numel()
            for key, buf in (("exp_avg", self._exp_avg), ("exp_avg_sq", self._exp_avg_sq)):
                view = buf[o:o + n].view(p.shape)
                if key in st and st[key].data_ptr() != view.data_ptr():
                    view.copy_(st[key])
                st[key] = view
            if "step" not in st:
                st["step"] = torch.tensor(0.0, dtype=torch.float32)
            step = max(step, int(st["step"].item()))
            self._step_tensors.append(st["step"])
        self._step_count = step
        self._bound_opt = opt

    def _still_bound(self, opt):
        """True while every live parameter's `exp_avg` / `exp_avg_sq` in `opt.state` is still the arena view bound earlier."""
        live = set(self.live_param_names())
        for name, p in self._arena_named_parameters():
            if name not in live:
                continue
            st = opt.state.get(p)
            o = self._offsets[name]
            if (not st or "exp_avg" not in st or st["exp_avg"].data_ptr() != self._exp_avg[o:].data_ptr()
                    or st["exp_avg_sq"].data_ptr() != self._exp_avg_sq[o:].data_ptr()):
                return False
        return all(a is b for a, b in zip(self._step_tensors, (opt.state[p]["step"] for n, p in self._arena_named_parameters() if n in live)))

    # -- CUDA graphs for launch-bound (small-batch) steps ---------------------------------------
    GRAPH_MAX_ROWS = 4096      # above this the step is GPU-bound and graphs buy nothing
    GRAPH_MAX_KEYS = 8         # captured graphs kept per model (least recently used dropped first)

    def graphed_step(self, key, inputs, body):
        """Run `body(static_inputs) -> dict of output tensors` as a replayed CUDA graph.

        A step at the reference's own batch sizes (128 / 256 rows) is ~70 short launches: host launch
        latency, not the GPU, sets the pace.  The first time a `key` (shapes + hyper-parameters) is seen the
        step runs eagerly (lazy one-time initialisation happens there), the second time it is captured, from
        then on it is replayed: inputs are copied into the graph's static buffers, outputs are read from them.
        """
        cache = self.__dict__.setdefault("_graphs", {})
        # A caller that feeds the SAME device tensors step after step (full-batch training on a resident table, the device-
        # timed bench) gets a graph that reads them in place: no copy into static buffers (at 2^22 tabular rows that copy was
        # a third of the step).  Detected, not requested: the pointers of two consecutive calls of one key agree -- the
        # previous call's tensors are kept alive until then, so an equal address is the same storage, never a recycled block.
        ptrs = tuple(v.data_ptr() for v in inputs.values() if v is not None)
        last = self.__dict__.setdefault("_graph_last_ptrs", {})
        if len(last) > 4 * self.GRAPH_MAX_KEYS:
            last.clear()
        alias = key in last and last[key][0] == ptrs
        last[key] = (ptrs, tuple(inputs.values()))
        if alias:
            key = key + ("in-place", ptrs)
        ent = cache.pop(key, None)
        if alias and ent is None:
            # the step has already run eagerly under the plain key (that is how `last` got its entry): capture right away
            while len(cache) >= self.GRAPH_MAX_KEYS:
                cache.pop(next(iter(cache)))
            ent = "seen"
        if ent is not None:
            cache[key] = ent                                   # most recently used last
        if ent is None:
            # the key holds hyper-parameters (lr, beta, lambda): a scheduler stepping them every iteration must not grow
            # one graph (static buffers + memory pool) per value -- keep the last few, and give graphs up when keys churn
            self._graph_misses = getattr(self, "_graph_misses", 0) + 1
            if self._graph_misses > self.GRAPH_MAX_KEYS * 4:
                self.use_graphs = False
                self.drop_graphs()
                return body(inputs)
            while len(cache) >= self.GRAPH_MAX_KEYS:
                cache.pop(next(iter(cache)))
            cache[key] = "seen"
            return body(inputs)
        if ent == "seen":
            if alias:
                static = dict(inputs)                          # the graph reads the caller's tensors (kept alive by this entry)
            else:
                static = {k: (None if v is None else torch.empty_like(v)) for k, v in inputs.items()}
                for k, v in inputs.items():
                    if v is not None:
                        static[k].copy_(v)
            if getattr(self, "_dev_step", None) is None:
                self._dev_step = torch.zeros(1, dtype=torch.int32, device=self.arena_device)
            self._dev_step.fill_(self._step_count)
            self._use_dev_step = True
            torch.cuda.synchronize(self.arena_device)
            graph = torch.cuda.CUDAGraph()
            count0, tensors0 = self._step_count, [float(t) for t in self._step_tensors[:1]]
            n0 = _lib.lib().cdg_launch_count()
            def _undo():
                # capturing does not execute: undo the host-side bookkeeping the body did (launch count, step counters)
                nodes = int(_lib.lib().cdg_launch_count() - n0)
                _lib.lib().cdg_launch_count_add(-nodes)
                undo = self._step_count - count0
                self._step_count = count0
                if undo:
                    torch._foreach_add_(self._step_tensors, -float(undo))
                return nodes
            try:
                with torch.cuda.graph(graph):
                    outs = body(static)
            except BaseException:
                _undo()
                cache.pop(key, None)
                raise
            finally:
                self._use_dev_step = False
            nodes = _undo()                                    # kernels of this library inside the graph
            ent = cache[key] = (graph, static, outs, nodes)
        graph, static, outs, nodes = ent
        if not alias:
            for k, v in inputs.items():
                if v is not None:
                    static[k].copy_(v, non_blocking=True)
        if int(self._dev_step_mirror) != self._step_count:
            self._dev_step.fill_(self._step_count)
        graph.replay()
        _lib.lib().cdg_launch_count_add(nodes)
        self._step_count += 1
        self._dev_step_mirror = self._step_count
        torch._foreach_add_(self._step_tensors, 1.0)
        return outs

    _dev_step_mirror = -1
    _use_dev_step = False

    def drop_graphs(self):
        self.__dict__["_graphs"] = {}
        self.__dict__["_graph_last_ptrs"] = {}

    def adam_step(self, grad_scale=1.0, clamp=None):
        g = self._opt_group
        a = _lib.AdamArgs()
        segs = self.adam_segments()
        if len(segs) > _lib.MAX_SEG:
            raise RuntimeError("too many Adam segments")
        a.n_seg = len(segs)
        for i, (o, n) in enumerate(segs):
            a.seg_off[i], a.seg_len[i] = o, n
        lr = g["lr"]
        a.lr = float(lr.item() if torch.is_tensor(lr) else lr)
        a.beta1, a.beta2 = float(g["betas"][0]), float(g["betas"][1])
        a.eps, a.weight_decay = float(g["eps"]), float(g["weight_decay"])
        a.grad_scale = float(grad_scale)
        self._step_count += 1
        a.step = self._step_count
        a.dev_step = _ptr(self._dev_step) if self._use_dev_step else None
        if clamp is not None:
            a.clamp_off, a.clamp_len, a.clamp_lo, a.clamp_hi = clamp
        else:
            a.clamp_off, a.clamp_len = -1, 0
        stream = torch.cuda.current_stream(self.arena_device).cuda_stream
        with torch.cuda.device(self.arena_device):
            _lib.check(_lib.lib().cdg_adam_step(_ptr(self._arena), _ptr(self._grads), _ptr(self._exp_avg),
                                                _ptr(self._exp_avg_sq), C.byref(a), C.c_void_p(stream)))
        torch._foreach_add_(self._step_tensors, 1.0)

    # -- the causal flows as the evaluation scripts call them (shared by every model family) -----
    @staticmethod
    def _cols(t):
        return list(torch.split(t, 1, dim=1))

    def _flows(self, x, direction, log_determinant=False):
        """All node flows on a [batch, node] matrix in one launch (parameters read in place from the arena)."""
        cfg = self.config
        return flow_apply(cfg["scm"], cfg.get("flow_num", 1), cfg.get("inverse_loop", 100), self._arena,
                          self._flow_offsets(cfg["node"]), x, direction, log_determinant)

    def _logdet(self, log_determinant, orig_latent):
        if not log_determinant:
            return [0] * self.config["node"]
        return self._cols(self._flows(orig_latent, 0, True)[1])                      # modules/model.py:22-25, :91-98

    def inverse(self, input):
        return self._cols(self._flows(torch.cat(list(input), dim=1), 1)[0])          # modules/model.py:252-254

    def transform(self, input, log_determinant=False):
        """u = input @ I_B_inv, then the per-node flows (modules/model.py:261-268).  The d x d product is bookkeeping on a
        [batch, d] matrix (torch.matmul on the device); the flows and their log|det| run in cdg_flow_apply."""
        orig_latent = torch.matmul(_f32c(input, self.arena_device), self.I_B_inv.to(self.arena_device))
        z, ld = self._flows(orig_latent, 0, log_determinant)
        return orig_latent, self._cols(z), (self._cols(ld) if log_determinant else [0] * self.config["node"])

    # -- data parallel -------------------------------------------------------------------------
    def exchange_gradients(self):
        """Sum the gradient arena over the data-parallel ranks.  Returns the factor the
        optimizer must scale gradients by (1/world)."""
        return _dist.allreduce_arena(self._grads, self.reduce_ranges())

    def reduce_ranges(self):
        return self.adam_segments()
