"""skeletondiffusion_b200 — B200-native (sm_100a) sampling path of SkeletonDiffusion.

Same Python API as the reference (`NonisotropicGaussianDiffusion`, `get_cov_from_corr`, `Denoiser`,
`AutoEncoder`, `get_prediction`), running hand-written CUDA kernels behind the C ABI declared in
include/skeldiff_b200.h.  There is no CPU fallback."""
from .diffusion import NonisotropicGaussianDiffusion, LatentDiffusion, get_cov_from_corr
from .network import Denoiser, StaticGraphLinear, Attention, ResnetBlock, Residual, PreNorm, RMSNorm
from .autoencoder import AutoEncoder, Encoder, Decoder, StaticGraphGRU
from .pipeline import (DiffusionManager, load_model_checkpoint, prepare_model, GraphedPrediction, best_sample, long_term_prediction_best_every50, get_prediction, get_diffusion_latent_codes, decode_latent_pred,
                       shard_windows, build_models)
from .skeletons import get_skeleton, SkeletonSpec
from .plan import invalidate_plans
from .metrics import motion_metrics, multimodal_metrics, ade, fde, apd, mmade, mmfde

__version__ = "0.1.0"
