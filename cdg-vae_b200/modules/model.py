"""Drop-in for the reference's modules/model.py (pendulum images): same class names, constructor
signatures, method arities, attribute names and state_dict keys; the arithmetic runs in
libcdgvae_sm100.so on an sm_100a device.

    CDGVAE(B, mask, config, device)               modules/model.py:208-304
    InvertiblePriorLinear(device)                  modules/model.py:8-29
    PlanarFlows(input_dim, flow_num, inverse_loop, device)   modules/model.py:31-100
"""
import ctypes as C

import torch
import torch.nn as nn

from .. import _lib
from ..engine import ArenaModule, _f32c, _ptr, flow_apply


class InvertiblePriorLinear(nn.Module):
    """Per-node affine flow o = p[0] * eps + p[1] (modules/model.py:8-29).  Holds the parameter; inside CDGVAE's step the
    evaluation is fused into the latent kernels, stand-alone calls (inference.py:317) run `cdg_flow_apply`."""

    def __init__(self, device="cpu"):
        super().__init__()
        self.p = nn.Parameter(torch.rand([2], device=device) * 0.1)        # model.py:18

    def forward(self, eps, log_determinant=False):
        o, ld = flow_apply("linear", 1, 0, self.p.data, [0], eps, 0, log_determinant)
        return o, (ld if log_determinant else 0)                            # model.py:20-25

    def inverse(self, o):
        return flow_apply("linear", 1, 0, self.p.data, [0], o, 1)[0]       # model.py:27-29


class PlanarFlows(nn.Module):
    """ELU planar flow with the invertibility re-parameterisation of u (modules/model.py:31-100), input_dim = 1 as every
    reference model builds it.  The arithmetic (build_u, the forward map with log|det|, the `inverse_loop` fixed-point
    inverse) runs in `cdg_flow_apply`; inside CDGVAE's step it is fused into the latent kernels."""

    def __init__(self, input_dim, flow_num, inverse_loop, device="cpu"):
        super().__init__()
        if input_dim != 1:
            raise ValueError("PlanarFlows: the reference only instantiates input_dim = 1 (modules/model.py:236); other widths are not built")
        self.input_dim, self.flow_num, self.inverse_loop, self.device = input_dim, flow_num, inverse_loop, device
        self.alpha = torch.tensor(1, dtype=torch.float32).to(device)
        self.w = nn.ParameterList([nn.Parameter(torch.randn(input_dim, 1, device=device) * 0.1) for _ in range(flow_num)])
        self.b = nn.ParameterList([nn.Parameter(torch.randn(1, 1, device=device) * 0.1) for _ in range(flow_num)])
        self.u = nn.ParameterList([nn.Parameter(torch.randn(input_dim, 1, device=device) * 0.1) for _ in range(flow_num)])

    def _packed(self):
        """{w[F], b[F], u[F]} as one contiguous fp32 block: the arena already lays a node's scalars out that way."""
        ts = [t.data for t in list(self.w) + list(self.b) + list(self.u)]
        base = ts[0]
        if all(t.dtype == torch.float32 and t.device == base.device and t.data_ptr() == base.data_ptr() + 4 * i for i, t in enumerate(ts)):
            return torch.as_strided(base, (len(ts),), (1,))
        return torch.cat([t.reshape(-1).float() for t in ts])

    def forward(self, inputs, log_determinant=False):
        o, ld = flow_apply("nonlinear", self.flow_num, self.inverse_loop, self._packed(), [0], inputs, 0, log_determinant)
        return o, (ld if log_determinant else 0)

    def inverse(self, inputs):
        return flow_apply("nonlinear", self.flow_num, self.inverse_loop, self._packed(), [0], inputs, 1)[0]


def mask_ranges(mask, n_cols):
    """Each decoder mask must be the {0,1} indicator of one contiguous range of flattened (H,W,3)
    columns (the row bands of main.py:167-179), pairwise disjoint.  Returns [(lo, hi)]."""
    out = []
    for k, m in enumerate(mask):
        flat = torch.as_tensor(m).reshape(-1).to("cpu")
        if flat.numel() != n_cols:
            raise ValueError(f"mask {k} has {flat.numel()} entries, expected {n_cols}")
        nz = torch.nonzero(flat).reshape(-1)
        if nz.numel() == 0:
            out.append((0, 0))
            continue
        lo, hi = int(nz[0]), int(nz[-1]) + 1
        if hi - lo != nz.numel() or not bool((flat[lo:hi] == 1).all()):
            raise ValueError(f"mask {k} is not the 0/1 indicator of a contiguous column range; "
                             "only band masks (main.py:167-179) are supported")
        out.append((lo, hi))
    return out


class CDGVAE(ArenaModule):
    HIDDEN = 300
    DEC_EXTRA_INPUTS = 0          # the DR variant appends the last latent to every decoder's input

    @staticmethod
    def _check_factor(config, mask):
        assert sum(config["factor"]) == config["node"]              # model.py:214
        assert len(config["factor"]) == len(mask)                   # model.py:215

    def __init__(self, B, mask, config, device):
        super().__init__()
        self.config = config
        self.mask = mask
        self._check_factor(config, mask)
        self.device = device
        P, H = 3 * config["image_size"] * config["image_size"], self.HIDDEN
        try:
            self._ranges = mask_ranges(mask, P)        # band masks (main.py:167-179): the fast path
            self._general_masks = None
        except ValueError:
            # anything else (overlapping, scattered, non-binary masks): every decoder computes all P columns and
            # xhat = tanh(sum_k out_k * mask_k) exactly as model.py:284-287 writes it
            self._ranges = [(0, P)] * len(mask)
            self._general_masks = torch.stack([torch.as_tensor(m, dtype=torch.float32).reshape(-1) for m in mask])

        # parameter creation order = reference order (encoder, flows, decoders) for same-seed init
        self.encoder = nn.Sequential(nn.Linear(P, H), nn.ELU(), nn.Linear(H, H), nn.ELU(),
                                     nn.Linear(H, config["node"] * 2)).to(device)
        self.B = B.to(device)
        self.I = torch.eye(config["node"]).to(device)
        self._A_host = torch.inverse(torch.eye(config["node"]) - B.detach().to("cpu", torch.float32))
        self.I_B_inv = self._A_host.to(device)                       # model.py:228-230
        if config["scm"] == "linear":
            self.flows = nn.ModuleList([InvertiblePriorLinear(device=device) for _ in range(config["node"])])
        elif config["scm"] == "nonlinear":
            self.flows = nn.ModuleList([PlanarFlows(1, config["flow_num"], config["inverse_loop"], device)
                                        for _ in range(config["node"])])
        else:
            raise ValueError("Not supported SCM!")                   # model.py:240
        self.decoder = nn.ModuleList([
            nn.Sequential(nn.Linear(k + self.DEC_EXTRA_INPUTS, H), nn.ELU(), nn.Linear(H, H), nn.ELU(), nn.Linear(H, P)).to(device)
            for k in config["factor"]])
        self.gemm_mode = config.get("gemm_mode", "auto") if isinstance(config, dict) else "auto"
        self.noise_fn = None      # tests/bench inject noise; default draws like model.py:276
        self._plan = None
        self._build_arena()

    # -- C-ABI plan ------------------------------------------------------------------------------
    def _destroy_plan(self):
        if getattr(self, "_plan", None):
            _lib.lib().cdg_pendulum_destroy(self._plan)
        self._plan = None

    def __del__(self):
        try:
            self._destroy_plan()
        except Exception:
            pass

    def _get_plan(self):
        cfg = self.config
        key = (float(cfg.get("beta", 0.0)), float(cfg.get("lambda", 0.0)), self.gemm_mode)
        if self._plan is not None and self._plan_key == key:
            return self._plan
        self._destroy_plan()
        _lib.require_cuda(self.arena_device)
        c = _lib.PendulumConfig()
        d, K = cfg["node"], len(cfg["factor"])
        if cfg["scm"] == "nonlinear" and cfg["flow_num"] > _lib.MAX_FLOW:
            raise ValueError(f"flow_num > {_lib.MAX_FLOW} is not supported")
        c.node, c.n_dec = d, K
        for k in range(K):
            c.factor[k] = cfg["factor"][k]
            c.dec_extra[k] = d - 1 if self.DEC_EXTRA_INPUTS else -1
            c.col_lo[k], c.col_hi[k] = self._ranges[k]
            for j, idx in enumerate((0, 2, 4)):
                c.dec[k][j] = self._lin(self._dec_prefix(k) + f".{idx}")
        c.scm, c.flow_num = _lib.SCM[cfg["scm"]], int(cfg.get("flow_num", 1))
        c.input_dim, c.hidden = 3 * cfg["image_size"] ** 2, self.HIDDEN
        c.gemm_mode = _lib.GEMM_MODES[self.gemm_mode]
        c.general_mask = int(getattr(self, "_general_masks", None) is not None)
        c.n_params = self._n_params
        for j, idx in enumerate((0, 2, 4)):
            c.enc[j] = self._lin(f"encoder.{idx}")
        for i, o in enumerate(self._flow_offsets(d)):
            c.flow_off[i] = o
        A = self._A_host.reshape(-1).tolist()
        for i, v in enumerate(A):
            c.I_B_inv[i] = v
        c.beta, c.lambda_ = key[0], key[1]
        plan = C.c_void_p()
        _lib.check(_lib.lib().cdg_pendulum_create(C.byref(c), C.byref(plan)))
        self._plan, self._plan_key = plan, key
        return plan

    def _dec_prefix(self, k):
        return f"decoder.{k}"

    def adam_segments(self):
        """Skip the decoder output rows the masks zero out: their gradient is exactly 0 in the reference,
        so Adam leaves them bit-unchanged."""
        segs, H = [], self.HIDDEN
        shapes = {n: p.numel() for n, p in self.named_parameters()}
        # with weight decay the reference's update of those rows is not zero (g = wd * p): then every row is live
        decays = bool(getattr(self, "_opt_group", None)) and float(self._opt_group.get("weight_decay", 0.0)) != 0.0
        for n in self.live_param_names():
            o = self._offsets[n]
            parts = n.split(".")
            if not decays and parts[0] == "decoder" and len(parts) == 4 and parts[2] == "4":
                lo, hi = self._ranges[int(parts[1])]
                w = H if parts[3] == "weight" else 1
                segs.append((o + lo * w, (hi - lo) * w))
            else:
                segs.append((o, shapes[n]))
        return self._merge(segs)

    # -- data parallel: all-reduce overlapped with the backward pass ------------------------------------
    def exchange_gradients(self):
        """All-reduce(sum) of the live gradient ranges, started bucket by bucket WHILE the backward pass is still running:
        the library records an event when a decoder's gradients are final (the decoders finish in index order, the encoder
        last), a communication stream waits for that event and issues the bucket's NCCL all-reduce, and the compute stream
        only waits for the collectives in front of the Adam kernel.  (The host is a whole step ahead of the device, so all of
        this is enqueued long before the first event fires.)  Returns the 1/world gradient scale."""
        from .. import dist as _dist
        import torch.distributed as tdist
        w = _dist.world()
        if w == 1:
            return 1.0
        plan = self._get_plan()
        lib = _lib.lib()
        dev = self.arena_device
        if getattr(self, "_ready_plan", None) is not plan:
            _lib.check(lib.cdg_pendulum_ready_events_enable(plan, 1))
            self._ready_plan = plan
            self._events_live = False            # the step that was just enqueued did not record them yet
        if not getattr(self, "_events_live", False):
            self._events_live = True
            return _dist.allreduce_arena(self._grads, self.reduce_ranges())
        side = self.__dict__.get("_comm_stream")
        if side is None:
            side = self.__dict__["_comm_stream"] = torch.cuda.Stream(device=dev)
        K = len(self.decoder) if isinstance(self.decoder, torch.nn.ModuleList) else 1
        # live ranges per bucket: decoder k's parameters are one contiguous block of the arena
        first = [self._offsets[n] for n in (f"decoder.{k}.0.weight" for k in range(K))] if isinstance(self.decoder, torch.nn.ModuleList) \
            else [self._offsets["decoder.0.weight"]]
        buckets = [[] for _ in range(K + 1)]
        for o, n in self.reduce_ranges():
            k = max([i for i in range(K) if o >= first[i]], default=K)
            buckets[k].append((o, n))
        handles = []
        for b in range(K + 1):
            ev = C.c_void_p()
            _lib.check(lib.cdg_pendulum_ready_event(plan, b, C.byref(ev)))
            _lib.check(lib.cdg_stream_wait_event(C.c_void_p(side.cuda_stream), ev))
            with torch.cuda.stream(side):
                for o, n in buckets[b]:
                    handles.append(tdist.all_reduce(self._grads[o:o + n], op=tdist.ReduceOp.SUM, async_op=True))
        for h in handles:
            h.wait()                              # the compute stream waits for the collectives (in front of Adam)
        return 1.0 / w

    def profile(self, enable=True):
        """Per-category device timing of the step (cudaEvents inside the library)."""
        _lib.check(_lib.lib().cdg_pendulum_profile_enable(self._get_plan(), int(enable)))
        self.use_graphs = not enable          # events recorded between launches cannot live inside a captured graph

    def profile_read(self):
        out = (C.c_double * len(_lib.PROF_CATS))()
        _lib.check(_lib.lib().cdg_pendulum_profile_read(self._get_plan(), out))
        return dict(zip(_lib.PROF_CATS, list(out)))

    def _masks_dev(self):
        m = getattr(self, "_general_masks", None)
        if m is None:
            return None
        if m.device != self.arena_device:
            self._general_masks = m = m.to(self.arena_device).contiguous()
        return m

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.arena_device).cuda_stream)

    def _noise(self, batch):
        if self.noise_fn is not None:
            return self.noise_fn(batch, self.config["node"])
        return torch.randn(batch, self.config["node"])              # CPU draw, as model.py:276

    # -- training entry (used by modules/train.py) -------------------------------------------------
    def forward_backward(self, x, y, noise, logs_row, x_l=None, y_l=None, xhat=None, infomax=None):
        """`infomax` = (discriminator, perm, gamma) runs the InfoMax discriminator inside the same step (train.py:113-142)."""
        dev = self.arena_device
        plan = self._get_plan()
        x = _f32c(x, dev).reshape(x.shape[0], -1)
        noise = _f32c(noise, dev)
        io = _lib.PendulumIO()
        io.params, io.grads = _ptr(self._arena), _ptr(self._grads)
        Bn = x.shape[0]
        keep = [x, noise]
        if x_l is not None:
            x_l = _f32c(x_l, dev).reshape(x_l.shape[0], -1)
            y_l = _f32c(y_l, dev)
            io.x_l, io.y_l, io.ld_y_l, io.batch_l = _ptr(x_l), _ptr(y_l), y_l.shape[1], x_l.shape[0]
            keep += [x_l, y_l]
        else:
            y = _f32c(y, dev)
            io.y, io.ld_y = _ptr(y), y.shape[1]
            keep.append(y)
        if infomax is not None:
            disc, perm, gamma = infomax
            perm = perm.to(device=dev, dtype=torch.int64).contiguous()
            io.d_params, io.d_grads, io.d_n_params = _ptr(disc._arena), _ptr(disc._grads), disc._n_params
            for j, idx in enumerate((0, 2, 4)):
                io.d_net[j] = disc._lin(f"net.{idx}")
            io.perm, io.gamma = _ptr(perm), float(gamma)
            keep.append(perm)
            nbytes = _lib.lib().cdg_pendulum_workspace_bytes_infomax(plan, Bn)
        else:
            nbytes = _lib.lib().cdg_pendulum_workspace_bytes(plan, Bn, 0 if x_l is None else x_l.shape[0])
        ws = self._get_workspace(nbytes)
        io.workspace, io.workspace_bytes = _ptr(ws), ws.numel()
        io.x, io.noise, io.batch = _ptr(x), _ptr(noise), Bn
        io.logs = _ptr(logs_row)
        io.xhat = _ptr(xhat)
        io.masks = _ptr(self._masks_dev())
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().cdg_pendulum_forward_backward(plan, C.byref(io), self._stream()))
        return keep

    # -- inference API (modules/model.py:252-304) ---------------------------------------------------
    def _run_forward(self, x=None, noise=None, latent_in=None, deterministic=False, want=()):
        dev = self.arena_device
        plan = self._get_plan()
        d = self.config["node"]
        P = 3 * self.config["image_size"] ** 2
        K = len(self.config["factor"])
        src = x if x is not None else latent_in
        Bn = src.shape[0]
        io = _lib.PendulumFwdIO()
        io.params = _ptr(self._arena)
        keep = []
        if x is not None:
            x = _f32c(x, dev).reshape(Bn, -1)
            io.x = _ptr(x)
            keep.append(x)
            if not deterministic:
                noise = _f32c(noise if noise is not None else self._noise(Bn), dev)
                io.noise = _ptr(noise)
                keep.append(noise)
        else:
            latent_in = _f32c(latent_in, dev)
            io.latent_in = _ptr(latent_in)
            keep.append(latent_in)
        out = {}
        for name in want:
            shape = {"xhat": (Bn, P), "xhat_separated": (K, Bn, P)}.get(name, (Bn, d))
            out[name] = torch.empty(shape, device=dev)
            setattr(io, name, _ptr(out[name]))
        nbytes = _lib.lib().cdg_pendulum_workspace_bytes(plan, Bn, 0)
        ws = self._get_workspace(nbytes)
        io.workspace, io.workspace_bytes, io.batch, io.deterministic = _ptr(ws), ws.numel(), Bn, int(deterministic)
        io.masks = _ptr(self._masks_dev())
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().cdg_pendulum_forward(plan, C.byref(io), self._stream()))
        return out

    def get_posterior(self, input):
        o = self._run_forward(x=input, deterministic=True, want=("mean", "logvar"))
        return o["mean"], o["logvar"]

    def encode(self, input, deterministic=False, log_determinant=False):
        o = self._run_forward(x=input, deterministic=deterministic,
                              want=("mean", "logvar", "epsilon", "orig_latent", "latent"))
        return (o["mean"], o["logvar"], o["epsilon"], o["orig_latent"], self._cols(o["latent"]),
                self._logdet(log_determinant, o["orig_latent"]))

    def decode(self, input):
        s = self.config["image_size"]
        o = self._run_forward(latent_in=torch.cat(list(input), dim=1), want=("xhat", "xhat_separated"))
        return list(o["xhat_separated"].unbind(0)), o["xhat"].view(-1, s, s, 3)

    def forward(self, input, deterministic=False, log_determinant=False):
        s = self.config["image_size"]
        o = self._run_forward(x=input, deterministic=deterministic,
                              want=("mean", "logvar", "epsilon", "orig_latent", "latent", "align_latent",
                                    "xhat_separated", "xhat"))
        return (o["mean"], o["logvar"], o["epsilon"], o["orig_latent"], self._cols(o["latent"]),
                self._logdet(log_determinant, o["orig_latent"]), self._cols(o["align_latent"]),
                list(o["xhat_separated"].unbind(0)), o["xhat"].view(-1, s, s, 3))


class VAE(CDGVAE):
    """The single-decoder baseline of modules/model.py:102-189: `VAE(B, config, device)`; forward returns the
    8-tuple (mean, logvar, epsilon, orig_latent, latent, logdet, align_latent, xhat).  It is the CDG-VAE step with one
    decoder that reads all `node` latents and owns every output column, so it runs on the same kernels."""

    def __init__(self, B, config, device):
        ArenaModule.__init__(self)
        self.config = config
        self.device = device
        P, H = 3 * config["image_size"] * config["image_size"], self.HIDDEN
        self.mask = [torch.ones(config["image_size"], config["image_size"], 3)]
        self._ranges = [(0, P)]
        self.encoder = nn.Sequential(nn.Linear(P, H), nn.ELU(), nn.Linear(H, H), nn.ELU(),
                                     nn.Linear(H, config["node"] * 2)).to(device)          # model.py:110-116
        self.B = B.to(device)
        self.I = torch.eye(config["node"]).to(device)
        self._A_host = torch.inverse(torch.eye(config["node"]) - B.detach().to("cpu", torch.float32))
        self.I_B_inv = self._A_host.to(device)
        if config["scm"] == "linear":
            self.flows = nn.ModuleList([InvertiblePriorLinear(device=device) for _ in range(config["node"])])
        elif config["scm"] == "nonlinear":
            self.flows = nn.ModuleList([PlanarFlows(1, config["flow_num"], config["inverse_loop"], device)
                                        for _ in range(config["node"])])
        else:
            raise ValueError("Not supported SCM!")
        self.decoder = nn.Sequential(nn.Linear(config["node"], H), nn.ELU(), nn.Linear(H, H), nn.ELU(), nn.Linear(H, P),
                                     nn.Tanh()).to(device)                                   # model.py:133-140
        self.gemm_mode = config.get("gemm_mode", "auto")
        self.noise_fn = None
        self._plan = None
        self._factor = [config["node"]]
        self._build_arena()

    def _dec_prefix(self, k):
        return "decoder"

    def _get_plan(self):
        # the C plan describes this model as one decoder over all latents
        cfg = dict(self.config)
        cfg["factor"] = self._factor
        saved, self.config = self.config, cfg
        try:
            return super()._get_plan()
        finally:
            self.config = saved

    def _run_forward(self, *a, **k):
        cfg = dict(self.config)
        cfg["factor"] = self._factor
        saved, self.config = self.config, cfg
        try:
            return super()._run_forward(*a, **k)
        finally:
            self.config = saved

    def decode(self, input):
        raise AttributeError("the reference VAE has no decode(); use forward()")

    def forward(self, input, deterministic=False, log_determinant=False):
        s = self.config["image_size"]
        o = self._run_forward(x=input, deterministic=deterministic,
                              want=("mean", "logvar", "epsilon", "orig_latent", "latent", "align_latent", "xhat"))
        return (o["mean"], o["logvar"], o["epsilon"], o["orig_latent"], self._cols(o["latent"]),
                self._logdet(log_determinant, o["orig_latent"]), self._cols(o["align_latent"]), o["xhat"].view(-1, s, s, 3))


class Discriminator(ArenaModule):
    """modules/model.py:191-206: the InfoMax critic on (x, z), Linear(P + node, 300)-ELU-Linear(300, 300)-ELU-Linear(300, 1).
    Inside train_InfoMax it is evaluated by the fused step (cdg_pendulum_forward_backward with d_params); `forward` here is
    the stand-alone evaluation on the parameter views."""

    def __init__(self, config, device="cpu"):
        super().__init__()
        self.config = config
        P = 3 * config["image_size"] * config["image_size"]
        self.net = nn.Sequential(nn.Linear(P + config["node"], 300), nn.ELU(), nn.Linear(300, 300), nn.ELU(),
                                 nn.Linear(300, 1)).to(device)
        self._build_arena()

    def forward(self, x, z):
        x = x.view(-1, 3 * self.config["image_size"] * self.config["image_size"])
        return self.net(torch.cat((x, z), dim=1))
