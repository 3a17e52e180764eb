"""DR/modules/train.py of the reference is byte-identical to modules/train.py: same loops."""
from ...modules.train import train_CDGVAE, train_CDGVAE_semi, train_CDGVAE_semi_loaders, train_VAE  # noqa: F401
