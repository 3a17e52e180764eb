"""Drop-in for the reference's tabular/modules/model.py: CDGVAE (loan / adult / covtype) and TVAE
(= CDG-TVAE).  Same constructor signatures, method arities, attributes and state_dict keys; the
arithmetic runs in libcdgvae_sm100.so (one fused kernel per step, csrc/tabular.cu).

    CDGVAE(B, mask, config, device)     tabular/modules/model.py:234-358
    TVAE(B, mask, config, device)       tabular/modules/model.py:360-460
    VAE(B, config, device)              tabular/modules/model.py:103-222   (baseline: one decoder over all latents)
    Discriminator(config, device)       tabular/modules/model.py:224-232   (InfoMax critic: parameter holder, see train_InfoMax)
"""
import ctypes as C

import torch
import torch.nn as nn

from ... import _lib
from ...engine import ArenaModule, _f32c, _ptr
from ...modules.model import InvertiblePriorLinear, PlanarFlows  # noqa: F401  (same classes as the pendulum tree)


class _TabularBase(ArenaModule):
    # The whole step is three short launches (step kernel, Adam, log row).  For loan / adult the step kernel runs ~0.12 ms
    # at 2^20 rows and the host-side launch sequence is as long again, so the step is replayed as a CUDA graph up to 2^22
    # rows (copying the batch into the graph's static buffers costs ~15 us at 2^20 rows): 4.4 -> 6.3 G rows/s.  covtype (0.24 ms per
    # 2^20 rows since the constant-parameter kernel) is replayed up to 2^21 rows; CDG-TVAE is GPU-bound from ~2^14 rows.
    @property
    def GRAPH_MAX_ROWS(self):
        if getattr(self, "sigma", None) is not None:
            return 1 << 14
        return (1 << 22) if self.config.get("dataset") in ("loan", "adult") else (1 << 21)

    KIND = None
    ENC_IDX = DEC_IDX = ()
    ACT = 0

    def _common_init(self, B, mask, config, device):
        self.config = config
        self.mask = mask
        assert sum(config["factor"]) == config["node"]               # model.py:240 / :366
        assert len(config["factor"]) == len(mask)                    # model.py:241 / :367
        self.device = device

    def _causal_init(self, B, config, device):
        self.B = B.to(device)
        self.I = torch.eye(config["node"]).to(device)
        self._A_host = torch.inverse(torch.eye(config["node"]) - B.detach().to("cpu", torch.float32))
        self.I_B_inv = self._A_host.to(device)                        # model.py:263-265
        if config["scm"] == "linear":
            self.flows = nn.ModuleList([InvertiblePriorLinear(device=device) for _ in range(config["node"])])
        elif config["scm"] == "nonlinear":
            self.flows = nn.ModuleList([PlanarFlows(1, config["flow_num"], config["inverse_loop"], device)
                                        for _ in range(config["node"])])
        else:
            raise ValueError("Not supported SCM!")                    # model.py:275

    # -- plan ------------------------------------------------------------------------------------
    def _destroy_plan(self):
        if getattr(self, "_plan", None):
            _lib.lib().cdg_tabular_destroy(self._plan)
        self._plan = None

    def __del__(self):
        try:
            self._destroy_plan()
        except Exception:
            pass

    def _kind(self):
        raise NotImplementedError

    def _get_plan(self, flatten_topology=None, output_info_list=None):
        cfg = self.config
        ft = tuple(flatten_topology) if flatten_topology is not None else None
        oil = None if output_info_list is None else tuple(tuple((int(s[0]), str(s[1])) for s in col) for col in output_info_list)
        beta = 1.0 if self.KIND == "tvae" else float(cfg.get("beta", 0.0))      # train_TVAE has no beta (train.py:307)
        key = (beta, float(cfg.get("lambda", 0.0)), ft, oil)
        if self._plan is not None and self._plan_key == key:
            return self._plan
        self._destroy_plan()
        _lib.require_cuda(self.arena_device)
        c = _lib.TabularConfig()
        d, K = cfg["node"], len(self.mask)
        c.kind = self._kind()
        c.node, c.n_dec = d, K
        for k in range(K):
            c.factor[k], c.out_dim[k] = cfg["factor"][k], int(self.mask[k])
            for j, idx in enumerate(self.DEC_IDX):
                c.dec[k][j] = self._lin(self._dec_name(k, idx))
        c.scm, c.flow_num = _lib.SCM[cfg["scm"]], int(cfg.get("flow_num", 1))
        if cfg["scm"] == "nonlinear" and c.flow_num > _lib.MAX_FLOW:
            raise ValueError(f"flow_num > {_lib.MAX_FLOW} is not supported")
        c.input_dim, c.act = cfg["input_dim"], self.ACT
        c.n_enc_layers, c.n_dec_layers = len(self.ENC_IDX), len(self.DEC_IDX)
        for j, idx in enumerate(self.ENC_IDX):
            c.enc[j] = self._lin(f"encoder.{idx}")
        for i, o in enumerate(self._flow_offsets(d)):
            c.flow_off[i] = o
        c.sigma_off = self._offsets.get("sigma", -1)
        if ft is not None:
            for j, v in enumerate(ft):
                c.flatten_topology[j] = int(v)
        c.n_span = 0
        if oil is not None:
            start, n = 0, 0
            for col in oil:
                for dim, fn in col:                                   # train.py:270-285: offsets advance by span dim
                    if n >= _lib.MAX_SPANS:
                        raise ValueError("too many output spans")
                    c.span_start[n], c.span_dim[n] = start, dim
                    c.span_kind[n] = 1 if fn == "softmax" else 0
                    start += dim
                    n += 1
            c.n_span = n
        c.n_params = self._n_params
        for i, v in enumerate(self._A_host.reshape(-1).tolist()):
            c.I_B_inv[i] = v
        c.beta, c.lambda_ = key[0], key[1]
        plan = C.c_void_p()
        _lib.check(_lib.lib().cdg_tabular_create(C.byref(c), C.byref(plan)))
        self._plan, self._plan_key = plan, key
        return plan

    def _dec_name(self, k, idx):
        return f"decoder.{k}.{idx}"

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.arena_device).cuda_stream)

    def _noise(self, batch):
        if self.noise_fn is not None:
            return self.noise_fn(batch, self.config["node"])
        return torch.randn(batch, self.config["node"])              # CPU draw, as model.py:327 / :429

    # -- training entry ------------------------------------------------------------------------------
    def forward_backward(self, x, y, noise, logs_row, flatten_topology=None, output_info_list=None):
        dev = self.arena_device
        plan = self._get_plan(flatten_topology, output_info_list)
        x = _f32c(x, dev).reshape(x.shape[0], -1)
        y, noise = _f32c(y, dev), _f32c(noise, dev)
        if y.shape[1] != self.config["node"]:
            raise ValueError("y must have `node` columns (tabular/modules/train.py:224 uses all of y_batch)")
        io = _lib.TabularIO()
        ws = self._get_workspace(256)
        io.params, io.grads, io.workspace, io.workspace_bytes = _ptr(self._arena), _ptr(self._grads), _ptr(ws), ws.numel()
        io.x, io.y, io.noise, io.batch, io.logs = _ptr(x), _ptr(y), _ptr(noise), x.shape[0], _ptr(logs_row)
        with torch.cuda.device(self.arena_device):
            _lib.check(_lib.lib().cdg_tabular_forward_backward(plan, C.byref(io), self._stream()))
        return [x, y, noise]

    # -- inference API ---------------------------------------------------------------------------------
    def _run_forward(self, x, noise=None, deterministic=False):
        dev = self.arena_device
        plan = self._get_plan(*getattr(self, "_last_aux", (self._default_ft(), None)))
        x = _f32c(x, dev).reshape(x.shape[0], -1)
        Bn, d = x.shape[0], self.config["node"]
        io = _lib.TabularIO()
        ws = self._get_workspace(256)
        io.params, io.workspace, io.workspace_bytes = _ptr(self._arena), _ptr(ws), ws.numel()
        io.x, io.batch = _ptr(x), Bn
        if not deterministic:
            noise = _f32c(noise if noise is not None else self._noise(Bn), dev)
            io.noise = _ptr(noise)
        xhat = torch.empty(Bn, int(sum(self.mask)), device=dev)
        lat = torch.empty(Bn, 6 * d, device=dev)
        io.xhat, io.latents = _ptr(xhat), _ptr(lat)
        with torch.cuda.device(self.arena_device):
            _lib.check(_lib.lib().cdg_tabular_forward(plan, C.byref(io), int(deterministic), self._stream()))
        return xhat, lat

    def _default_ft(self):
        return None

    def get_posterior(self, input):
        d = self.config["node"]
        _, lat = self._run_forward(input, deterministic=True)
        return lat[:, :d].contiguous(), lat[:, d:2 * d].contiguous()

    def _unpack(self, lat):
        d = self.config["node"]
        return [lat[:, i * d:(i + 1) * d].contiguous() for i in range(6)]

    def encode(self, input, deterministic=False, log_determinant=False):
        _, lat = self._run_forward(input, deterministic=deterministic)
        mean, logvar, eps, orig, z, _ = self._unpack(lat)
        return mean, logvar, eps, orig, self._cols(z), self._logdet(log_determinant, orig)

    def decode(self, input):
        """xhat_separated, xhat = cat (model.py:337-342).  Off the hot path: evaluated with the parameter views."""
        latent = torch.split(torch.cat(list(input), dim=1), self.config["factor"], dim=-1)
        sep = [D(z) for D, z in zip(self.decoder, latent)]
        return sep, torch.cat(sep, dim=1)

    def forward(self, input, deterministic=False, log_determinant=False):
        xhat, lat = self._run_forward(input, deterministic=deterministic)
        mean, logvar, eps, orig, z, zal = self._unpack(lat)
        sep = list(torch.split(xhat, [int(m) for m in self.mask], dim=1))
        return (mean, logvar, eps, orig, self._cols(z), self._logdet(log_determinant, orig), self._cols(zal),
                sep, xhat)


class CDGVAE(_TabularBase):
    KIND = "cdgvae"

    def __init__(self, B, mask, config, device):
        super().__init__()
        self._common_init(B, mask, config, device)
        cov = config["dataset"] == "covtype"
        d = config["node"]
        if cov:                                                          # model.py:245-254
            self.encoder = nn.Sequential(nn.Linear(config["input_dim"], 4), nn.ELU(), nn.Linear(4, 4), nn.ELU(),
                                         nn.Linear(4, 4), nn.ELU(), nn.Linear(4, d * 2)).to(device)
        else:                                                            # model.py:256-260
            self.encoder = nn.Sequential(nn.Linear(config["input_dim"], 4), nn.ELU(), nn.Linear(4, d * 2)).to(device)
        self._causal_init(B, config, device)
        if cov:                                                          # model.py:278-298 (incl. the unused 7th decoder)
            net = [nn.Sequential(nn.Linear(k, 2), nn.ELU(), nn.Linear(2, 2), nn.ELU(), nn.Linear(2, m)).to(device)
                   for k, m in zip(config["factor"], self.mask)]
            net += [nn.Sequential(nn.Linear(config["factor"][-1], 4), nn.ELU(), nn.Linear(4, 4), nn.ELU(),
                                  nn.Linear(4, 8), nn.ELU(), nn.Linear(8, self.mask[-1])).to(device)]
            self.decoder = nn.ModuleList(net)
        else:                                                            # model.py:300-305
            self.decoder = nn.ModuleList([nn.Sequential(nn.Linear(k, 2), nn.ELU(), nn.Linear(2, m)).to(device)
                                          for k, m in zip(config["factor"], self.mask)])
        self.ENC_IDX = (0, 2, 4, 6) if cov else (0, 2)
        self.DEC_IDX = (0, 2, 4) if cov else (0, 2)
        self.ACT = 0
        self.noise_fn = None
        self._plan = None
        self._build_arena()

    def _kind(self):
        ds = self.config["dataset"]
        if ds not in ("loan", "adult", "covtype"):
            raise ValueError("Not supported dataset!")                   # train.py:210
        return _lib.TAB_KIND[ds]

    def _default_ft(self):
        return {"loan": (1, 2, 3, 4, 0), "adult": (2, 3, 0, 1, 4)}.get(self.config["dataset"])

    def live_param_names(self):
        # covtype: decoder.6.* exists in the state_dict but never receives a gradient (model.py:340 zip)
        n_used = len(self.mask)
        return [n for n, _ in self.named_parameters()
                if not (n.startswith("decoder.") and int(n.split(".")[1]) >= n_used)]


class VAE(_TabularBase):
    """tabular/modules/model.py:103-222: the CDG-VAE encoder and causal layer with ONE decoder over all latents.  The fused
    step kernel takes it as a single-factor model (factor = [node], out_dim = [D']); layer widths other than the CDG-VAE's go
    through the generic per-row kernel (csrc/tabular.cu)."""
    KIND = "vae"

    def __init__(self, B, config, device):
        super().__init__()
        d, D = config["node"], config["input_dim"]
        ds = config["dataset"]
        self.config = dict(config, factor=[d])
        self.device = device
        cov = ds == "covtype"
        out = D - 1 + 7 if cov else D
        self.mask = [out]
        if cov:                                                          # model.py:111-119
            self.encoder = nn.Sequential(nn.Linear(D, 4), nn.ELU(), nn.Linear(4, 4), nn.ELU(), nn.Linear(4, 4), nn.ELU(),
                                         nn.Linear(4, d * 2)).to(device)
        else:                                                            # model.py:121-125
            self.encoder = nn.Sequential(nn.Linear(D, 4), nn.ELU(), nn.Linear(4, d * 2)).to(device)
        self._causal_init(B, config, device)
        if ds == "loan":                                                 # model.py:144-148
            self.decoder = nn.Sequential(nn.Linear(d, 4), nn.ELU(), nn.Linear(4, out)).to(device)
        elif ds in ("adult", "covtype"):                                 # model.py:149-169
            self.decoder = nn.Sequential(nn.Linear(d, 8), nn.ELU(), nn.Linear(8, 8), nn.ELU(), nn.Linear(8, 16), nn.ELU(),
                                         nn.Linear(16, out)).to(device)
        else:
            raise ValueError("Not supported dataset!")                   # model.py:171
        self.ENC_IDX = (0, 2, 4, 6) if cov else (0, 2)
        self.DEC_IDX = (0, 2) if ds == "loan" else (0, 2, 4, 6)
        self.ACT = 0
        self.noise_fn = None
        self._plan = None
        self._build_arena()

    def _dec_name(self, k, idx):
        return f"decoder.{idx}"

    def _kind(self):
        return _lib.TAB_KIND[self.config["dataset"]]

    def _default_ft(self):
        return {"loan": (1, 2, 3, 4, 0), "adult": (2, 3, 0, 1, 4)}.get(self.config["dataset"])

    def live_param_names(self):
        return [n for n, _ in self.named_parameters()]

    def decode(self, input):
        xhat = self.decoder(torch.cat(list(input), dim=1))
        return [xhat], xhat

    def forward(self, input, deterministic=False, log_determinant=False):
        """8-tuple, no xhat_separated (model.py:222)."""
        xhat, lat = self._run_forward(input, deterministic=deterministic)
        mean, logvar, eps, orig, z, zal = self._unpack(lat)
        return (mean, logvar, eps, orig, self._cols(z), self._logdet(log_determinant, orig), self._cols(zal), xhat)


class Discriminator(nn.Module):
    """tabular/modules/model.py:224-232: the InfoMax critic on (x, epsilon).  A plain parameter holder with the reference's
    module tree (`net.0`, `net.2`), so checkpoints load; its training step (tabular train_InfoMax) is not built -- see there."""

    def __init__(self, config, device="cpu"):
        super().__init__()
        self.config = config
        self.net = nn.Sequential(nn.Linear(config["input_dim"] + config["node"], 4), nn.ELU(), nn.Linear(4, 1)).to(device)

    def forward(self, x, z):
        x = x.view(-1, self.config["input_dim"])
        return self.net(torch.cat((x, z), dim=1))


class TVAE(_TabularBase):
    KIND = "tvae"

    def __init__(self, B, mask, config, device):
        super().__init__()
        self._common_init(B, mask, config, device)
        d = config["node"]
        self.encoder = nn.Sequential(nn.Linear(config["input_dim"], 32), nn.ReLU(), nn.Linear(32, 16), nn.ReLU(),
                                     nn.Linear(16, 16), nn.ReLU(), nn.Linear(16, d * 2)).to(device)     # model.py:371-379
        self._causal_init(B, config, device)
        self.decoder = nn.ModuleList([
            nn.Sequential(nn.Linear(k, 8), nn.ReLU(), nn.Linear(8, 8), nn.ReLU(), nn.Linear(8, 16), nn.ReLU(),
                          nn.Linear(16, m)).to(device) for k, m in zip(config["factor"], self.mask)])   # model.py:397-406
        self.sigma = nn.Parameter(torch.ones(config["input_dim"]) * 0.1)                               # model.py:407
        self.ENC_IDX = (0, 2, 4, 6)
        self.DEC_IDX = (0, 2, 4, 6)
        self.ACT = 1
        self.noise_fn = None
        self._plan = None
        self._build_arena()

    def _kind(self):
        return _lib.TAB_KIND["tvae"]

    def _run_forward(self, x, noise=None, deterministic=False):
        if getattr(self, "_last_aux", None) is None:
            # outside train_TVAE no span table is needed for a forward pass: one softmax span over everything
            self._last_aux = (None, [[(int(sum(self.mask)), "softmax")]])
        return super()._run_forward(x, noise, deterministic)
