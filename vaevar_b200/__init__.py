"""vaevar_b200 -- B200-native 4D-Var cost-and-gradient engine behind the call surface of xiaoyi018/VAE-Var's
da_4dvar.py (one_step_DA[vae4dvar]), nf_model/vae.py (VAE_lr) and networks_old/transformer.py (LGUnet_all)."""
from .config import DECODER_FULL, ENCODER_FULL, FLOW_FULL, NetConfig, era5_stats, small  # noqa: F401
