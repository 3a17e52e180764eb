"""Drop-in for the reference's celeba/module/model.py: same class names, constructor signature, method arities,
attribute names and state_dict keys; the arithmetic runs in libcdgvae_sm100.so (csrc/celeba_step.cu).

    CDGVAE(B, mask, config, device, fc_size=32)        celeba/module/model.py:106-218

Reference behaviours kept: `self.decoder` is a plain Python list, so the five generators are not
sub-modules (absent from parameters() / state_dict(), never optimised, always in training mode); the ResNet-18
encoder is frozen except its new `fc` and runs with batch statistics; every forward advances the BatchNorm running
statistics (twice per forward() for the encoder: two encode() calls) and the spectral-norm u / v vectors.

Storage: the trainable parameters (encoder.fc, flows) live in the ArenaModule arena, everything else — frozen
encoder weights, BatchNorm buffers, generator weights and buffers — in one flat "frozen" arena the kernels index.
"""
import ctypes as C
import warnings

import torch
import torch.nn as nn

from ... import _lib
from ...engine import ALIGN, ArenaModule, _f32c, _ptr
from ...modules.model import InvertiblePriorLinear, PlanarFlows  # noqa: F401  (same classes as celeba/module/model.py:10-104)
from .sagan import Generator

# latent columns each decoder reads (model.py:190-194); the fifth reads epsilon2 (model.py:195)
DECODER_INPUTS = [[0, 2], [0, 3], [0, 4], [0, 1, 5]]
LOG_KEYS = ["loss", "recon", "KL", "alignment", "active"]


def _resnet18(pretrained):
    from torchvision import models
    if pretrained:
        try:
            return models.resnet18(weights=models.ResNet18_Weights.IMAGENET1K_V1)      # model.py:117 pretrained=True
        except Exception as e:      # no network / no cached checkpoint
            warnings.warn(f"ImageNet weights for resnet18 are unavailable ({e}); the encoder starts from random weights")
    return models.resnet18(weights=None)


class CDGVAE(ArenaModule):
    IMAGE_SIZE = 128

    def __init__(self, B, mask, config, device, fc_size=32):
        super().__init__()
        self.config = config
        self.mask = mask
        self.device = device
        self.encoder = _resnet18(config.get("pretrained", True))
        self.encoder.fc = nn.Linear(self.encoder.fc.in_features, config["node"] * 2 + config["latent_dim"] * 2)   # model.py:118
        for p in self.encoder.parameters():                                            # model.py:121-125
            p.requires_grad_(False)
        self.encoder.fc.weight.requires_grad = True
        self.encoder.fc.bias.requires_grad = True
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
        # a plain list, as in the reference (model.py:146-151)
        self.decoder = [Generator(2).to(device), Generator(2).to(device), Generator(2).to(device), Generator(3).to(device),
                        Generator(config["latent_dim"]).to(device)]
        self.gemm_mode = config.get("gemm_mode", "auto")
        self.noise_fn = None
        self._plan = None
        self._masks_cache = None
        self.train()
        self.to(device)

    # -- arenas ----------------------------------------------------------------------------------------------
    def _arena_named_parameters(self):
        return [(n, p) for n, p in self.named_parameters() if p.requires_grad]

    def _frozen_named_tensors(self):
        out = [(n, p) for n, p in self.named_parameters() if not p.requires_grad]
        out += [(n, b) for n, b in self.named_buffers() if b.dtype.is_floating_point]
        for k, g in enumerate(self.decoder):
            out += [(f"decoder.{k}.{n}", p) for n, p in g.named_parameters()]
            out += [(f"decoder.{k}.{n}", b) for n, b in g.named_buffers() if b.dtype.is_floating_point]
        return out

    def _build_arena(self):
        if not hasattr(self, "decoder"):
            return
        super()._build_arena()
        device = self._arena.device
        named = self._frozen_named_tensors()
        off, offs = 0, {}
        for n, t in named:
            offs[n] = off
            off += (t.numel() + ALIGN - 1) // ALIGN * ALIGN
        frozen = torch.zeros(off, dtype=torch.float32, device=device)
        with torch.no_grad():
            for n, t in named:
                o, k = offs[n], t.numel()
                frozen[o:o + k].copy_(t.detach().reshape(-1).to(torch.float32))
                t.data = frozen[o:o + k].view(t.shape)
        for g in self.decoder:                       # integer buffers (num_batches_tracked) follow the arena's device
            for b in g.buffers():
                if not b.dtype.is_floating_point:
                    b.data = b.data.to(device)
        self._frozen, self._foff = frozen, offs
        self._masks_cache = None

    def _destroy_plan(self):
        if getattr(self, "_plan", None):
            _lib.lib().cdg_celeba_destroy(self._plan)
        self._plan = None

    def __del__(self):
        try:
            self._destroy_plan()
        except Exception:
            pass

    # -- C-ABI plan --------------------------------------------------------------------------------------------
    def _conv(self, prefix, mod, sn=False, linear=False):
        c = _lib.Conv()
        w = mod.weight_orig if sn else mod.weight
        c.w = self._foff[prefix + (".weight_orig" if sn else ".weight")]
        c.b = self._foff[prefix + ".bias"] if mod.bias is not None else -1
        c.u = self._foff[prefix + ".weight_u"] if sn else -1
        c.v = self._foff[prefix + ".weight_v"] if sn else -1
        c.cout, c.cin = w.shape[0], w.shape[1]
        c.k, c.stride, c.pad = (1, 1, 0) if linear else (mod.kernel_size[0], mod.stride[0], mod.padding[0])
        return c

    def _bn(self, prefix, mod):
        b = _lib.BNorm()
        b.weight, b.bias = self._foff[prefix + ".weight"], self._foff[prefix + ".bias"]
        b.running_mean, b.running_var = self._foff[prefix + ".running_mean"], self._foff[prefix + ".running_var"]
        b.c, b.momentum, b.eps = mod.num_features, mod.momentum, mod.eps
        return b

    def _get_plan(self):
        cfg = self.config
        key = (float(cfg.get("beta", 0.0)), float(cfg.get("lambda", 0.0)), self.gemm_mode)
        if self._plan is not None and self._plan_key == key:
            return self._plan
        self._destroy_plan()
        _lib.require_cuda(self.arena_device)
        c = _lib.CelebaConfig()
        d = cfg["node"]
        c.node, c.latent_dim = d, cfg["latent_dim"]
        c.scm, c.flow_num = _lib.SCM[cfg["scm"]], int(cfg.get("flow_num", 1))
        c.image_size, c.gemm_mode = self.IMAGE_SIZE, _lib.GEMM_MODES[self.gemm_mode]
        c.n_params, c.n_frozen = self._n_params, self._frozen.numel()
        c.fc = self._lin("encoder.fc")
        for i, o in enumerate(self._flow_offsets(d)):
            c.flow_off[i] = o
        for i, v in enumerate(self._A_host.reshape(-1).tolist()):
            c.I_B_inv[i] = v
        c.beta, c.lambda_ = key[0], key[1]
        e = self.encoder
        c.rn_conv1, c.rn_bn1 = self._conv("encoder.conv1", e.conv1), self._bn("encoder.bn1", e.bn1)
        i = 0
        for li in range(1, 5):
            for bi, blk in enumerate(getattr(e, f"layer{li}")):
                p = f"encoder.layer{li}.{bi}"
                rb = c.rn_blk[i]
                rb.conv1, rb.conv2 = self._conv(p + ".conv1", blk.conv1), self._conv(p + ".conv2", blk.conv2)
                rb.bn1, rb.bn2 = self._bn(p + ".bn1", blk.bn1), self._bn(p + ".bn2", blk.bn2)
                rb.has_down = int(blk.downsample is not None)
                if blk.downsample is not None:
                    rb.down, rb.bn_down = self._conv(p + ".downsample.0", blk.downsample[0]), self._bn(p + ".downsample.1", blk.downsample[1])
                i += 1
        for k, g in enumerate(self.decoder):
            G = c.gen[k]
            src = DECODER_INPUTS[k] if k < len(DECODER_INPUTS) else [-(j + 1) for j in range(cfg["latent_dim"])]
            G.z_dim = len(src)
            for j, s in enumerate(src):
                G.z_src[j] = s
            pre = f"decoder.{k}."
            G.lin0 = self._conv(pre + "block0.snlinear0", g.block0.snlinear0, sn=True, linear=True)
            for b in range(5):
                blk = getattr(g, f"block{b + 1}")
                q = f"{pre}block{b + 1}."
                G.blk[b].conv1 = self._conv(q + "conv_1", blk.conv_1, sn=True)
                G.blk[b].conv2 = self._conv(q + "conv_2", blk.conv_2, sn=True)
                G.blk[b].conv0 = self._conv(q + "conv_0", blk.conv_0, sn=True)
                G.blk[b].bn1, G.blk[b].bn2 = self._bn(q + "bn1", blk.bn1), self._bn(q + "bn2", blk.bn2)
            for j, nm in enumerate(("theta", "phi", "g", "attn")):
                G.attn[j] = self._conv(f"{pre}self_attn1.snconv1x1_{nm}", getattr(g.self_attn1, f"snconv1x1_{nm}"), sn=True)
            G.bn = self._bn(pre + "bn", g.bn)
            G.to_rgb = self._conv(pre + "toRGB", g.toRGB, sn=True)
        plan = C.c_void_p()
        _lib.check(_lib.lib().cdg_celeba_create(C.byref(c), C.byref(plan)))
        self._plan, self._plan_key = plan, key
        return plan

    # -- helpers ---------------------------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.arena_device).cuda_stream)

    def _masks_dev(self, batch):
        m = self._masks_cache
        if m is None:
            m = torch.stack([torch.as_tensor(t, dtype=torch.float32).reshape(t.shape[0], -1) for t in self.mask])
            self._masks_cache = m = m.to(self.arena_device).contiguous()          # [5, B, S*S]
        if m.shape[1] != batch:
            raise ValueError(f"the decoder masks were taken from a batch of {m.shape[1]} (celeba/main.py:111); got a batch of {batch}")
        return m

    def _noise(self, batch):
        if self.noise_fn is not None:
            return self.noise_fn(batch, self.config["node"])
        return torch.randn(batch, self.config["node"])                       # CPU draws, as model.py:182 / :184

    def _bump_counters(self, encoder_passes, generators):
        """num_batches_tracked of the BatchNorms that ran (integer buffers, host-side bookkeeping)."""
        if encoder_passes:
            torch._foreach_add_([b for n, b in self.encoder.named_buffers() if n.endswith("num_batches_tracked")], encoder_passes)
        if generators:
            for g in self.decoder:
                torch._foreach_add_([b for n, b in g.named_buffers() if n.endswith("num_batches_tracked")], 1)

    def _run(self, x=None, y=None, noise=None, backward=False, deterministic=False, encoder_passes=2, encode_only=False,
             latent_in=None, epsilon2_in=None, logs_row=None, want_xhat=False, want_sep=False, want_latents=False):
        if latent_in is None and not self.encoder.training:
            # the reference only ever runs this model in training mode (celeba/main.py:118); eval-mode BatchNorm in the
            # encoder (running statistics instead of batch statistics) is not implemented in the kernels
            raise NotImplementedError("the CelebA encoder is evaluated with batch statistics (model.train()); eval mode is not built")
        dev = self.arena_device
        plan = self._get_plan()
        d, S = self.config["node"], self.IMAGE_SIZE
        io = _lib.CelebaIO()
        keep = []
        if latent_in is not None:
            latent_in, epsilon2_in = _f32c(latent_in, dev), _f32c(epsilon2_in, dev)
            Bn = latent_in.shape[0]
            io.latent_in, io.epsilon2_in = _ptr(latent_in), _ptr(epsilon2_in)
            keep += [latent_in, epsilon2_in]
            encoder_passes = 0
        else:
            x = _f32c(x, dev)
            Bn = x.shape[0]
            if tuple(x.shape[1:3]) != (S, S) or x.shape[3] < 3:
                raise ValueError(f"expected images of shape [B, {S}, {S}, >=3], got {tuple(x.shape)}")
            io.x, io.ld_x = _ptr(x), x.shape[3]
            keep.append(x)
            if not encode_only:
                self._masks_dev(Bn)                 # batch-shaped masks (celeba/main.py:111): fail before any work
            if not deterministic:
                n1, n2 = noise if noise is not None else (self._noise(Bn), self._noise(Bn))
                n1, n2 = _f32c(n1, dev), _f32c(n2, dev)
                io.noise1, io.noise2 = _ptr(n1), _ptr(n2)
                keep += [n1, n2]
        if y is not None:
            y = _f32c(y, dev)
            io.y, io.ld_y = _ptr(y), y.shape[1]
            keep.append(y)
        out = {}
        if not encode_only:
            io.masks = _ptr(self._masks_dev(Bn))
            if want_xhat:
                out["xhat"] = torch.empty(Bn, S, S, 3, device=dev)
                io.xhat = _ptr(out["xhat"])
            if want_sep:
                out["sep"] = torch.empty(5, Bn, S, S, 3, device=dev)
                io.xhat_separated = _ptr(out["sep"])
        if want_latents:
            out["latents"] = torch.empty(9, Bn, d, device=dev)
            io.latents = _ptr(out["latents"])
        io.params, io.grads, io.frozen = _ptr(self._arena), _ptr(self._grads), _ptr(self._frozen)
        nbytes = _lib.lib().cdg_celeba_workspace_bytes(plan, Bn)
        if nbytes < 0:
            _lib.check(1)
        ws = self._get_workspace(nbytes)
        io.workspace, io.workspace_bytes, io.batch = _ptr(ws), ws.numel(), Bn
        io.backward, io.deterministic, io.encoder_passes, io.encode_only = int(backward), int(deterministic), int(encoder_passes), int(encode_only)
        io.logs = _ptr(logs_row)
        with torch.cuda.device(self.arena_device):
            _lib.check(_lib.lib().cdg_celeba_step(plan, C.byref(io), self._stream()))
        self._bump_counters(encoder_passes, generators=not encode_only)
        out["keep"] = keep
        return out

    # -- training entry (celeba/module/train.py) ---------------------------------------------------------------------
    def forward_backward(self, x, y, noise, logs_row, xhat=False):
        return self._run(x=x, y=y, noise=noise, backward=True, logs_row=logs_row, want_xhat=xhat)

    def live_param_names(self):
        return [n for n, _ in self._arena_named_parameters()]

    # -- the reference's public methods -------------------------------------------------------------------------------
    def get_posterior(self, input):
        L = self._run(x=input, deterministic=True, encoder_passes=1, encode_only=True, want_latents=True)["latents"]
        return L[0], L[1], L[6], L[7]

    def encode(self, input, deterministic=False, log_determinant=False):
        L = self._run(x=input, deterministic=deterministic, encoder_passes=1, encode_only=True, want_latents=True)["latents"]
        return ((L[0], L[1], L[2], L[3], self._cols(L[4]), self._logdet(log_determinant, L[3])), (L[6], L[7], L[8]))

    def decode(self, latent, epsilon2):
        o = self._run(latent_in=torch.cat(list(latent), dim=1), epsilon2_in=epsilon2, want_xhat=True, want_sep=True)
        return [s.permute(0, 3, 1, 2) for s in o["sep"].unbind(0)], o["xhat"]       # generators emit NCHW (model.py:196)

    def forward(self, input, deterministic=False, log_determinant=False):
        o = self._run(x=input, deterministic=deterministic, encoder_passes=2, want_xhat=True, want_sep=True, want_latents=True)
        L = o["latents"]
        n = input.shape[0]
        return ((L[0], L[1], L[2], L[3], self._cols(L[4]), self._logdet(log_determinant, L[3])), (L[6], L[7], L[8]),
                self._cols(L[5]), [s.permute(0, 3, 1, 2) for s in o["sep"].unbind(0)], o["xhat"])
