"""Host-side mirrors of the reference's model/ package for the denoise hot path (same class names,
forward signatures and state-dict semantics; arithmetic in edgestyle_b200.engine)."""
from .controllora import (CachedControlNetModel, ControlLoRAModel, ControlNetOutput,  # noqa: F401
                          FusedControlLoRAModel, UNet2DConditionModel)
from .edgestyle_multicontrolnet import EdgeStyleMultiControlNetModel  # noqa: F401
from .edgestyle_pipeline import EdgeStyleStableDiffusionControlNetPipeline, StableDiffusionPipelineOutput  # noqa: F401
from ..vae import AutoencoderKL  # noqa: F401  (diffusers AutoencoderKL surface used by controllora.py:38-42 and edgestyle_pipeline.py:552-557)
