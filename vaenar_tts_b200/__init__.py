"""vaenar_tts_b200 -- B200-native (sm_100a) implementation of the VAENAR-TTS mel-synthesis hot path.

``VAENAR`` mirrors models.models.VAENAR of the reference; all compute runs in libvaenar_sm100.so
(hand-written CUDA: tcgen05/TMEM/TMA GEMM + attention kernels, fp32 flow kernels)."""
from .hparams import LJHPS, DataBakerHPS  # noqa: F401
from .model import VAENAR, InferenceSession  # noqa: F401
from ._lib import VaenarError, LIB_PATH, EXPORTS  # noqa: F401
