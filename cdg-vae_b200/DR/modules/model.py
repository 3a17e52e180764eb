"""Drop-in for the reference's DR/modules/model.py (distributional-robustness variant of the pendulum model):
`CDGVAE(B, mask, config, device)` with `node = sum(factor) + 1`; every decoder's first Linear takes its factor's
latents plus the last ("spurious") latent, `nn.Linear(k+1, 300)` (DR/modules/model.py:245, :284-287).  Everything
else — encoder, flows, masks, losses, the train loops (DR/modules/train.py is byte-identical to modules/train.py) —
is the pendulum path, so it runs on the same kernels."""
from ...modules.model import CDGVAE as _CDGVAE, VAE, InvertiblePriorLinear, PlanarFlows  # noqa: F401


class CDGVAE(_CDGVAE):
    DEC_EXTRA_INPUTS = 1

    @staticmethod
    def _check_factor(config, mask):
        # the reference comments the sum(factor) == node assert out here (DR/modules/model.py:214)
        assert len(config["factor"]) == len(mask)
        if sum(config["factor"]) != config["node"] - 1:
            raise ValueError("DR CDGVAE expects sum(factor) == node - 1 (the last latent is the spurious one)")

    def decode(self, input):
        import torch
        s = self.config["image_size"]
        o = self._run_forward(latent_in=torch.cat(list(input), dim=1), want=("xhat", "xhat_separated"))
        return list(o["xhat_separated"].unbind(0)), o["xhat"].view(-1, s, s, 3)
