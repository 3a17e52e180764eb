"""cdgvae_b200 — B200 (sm_100a) implementation of the CDG-VAE / CDG-TVAE training step behind the
reference's own Python API.

    from cdgvae_b200.modules.model import CDGVAE              # pendulum   (reference modules/model.py)
    from cdgvae_b200.modules.train import train_CDGVAE, train_CDGVAE_semi
    from cdgvae_b200.tabular.modules.model import CDGVAE, TVAE   # tabular  (reference tabular/modules/model.py)
    from cdgvae_b200.tabular.modules.train import train_CDGVAE, train_TVAE

Everything numerical runs in libcdgvae_sm100.so (include/cdgvae.h); there is no CPU fallback.
"""
from . import _lib, build, dist  # noqa: F401

__all__ = ["_lib", "build", "dist"]
