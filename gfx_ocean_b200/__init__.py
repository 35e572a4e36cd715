"""gfx_ocean_b200 -- B200-native replacement for gfx-ocean's per-frame compute path
(propagate -> 2-D inverse FFT of height/dx/dz -> sign correction), behind the C ABI of
include/ocean_b200.h. See DESIGN.md."""
from ._lib import (OceanError, PropagateLocals, CorrectionLocals, SpectrumParams, PIPELINE_FUSED, PIPELINE_LITERAL,  # noqa: F401
                   FLAG_DOUBLE_BUFFER_OUTPUT, FLAG_DX_PLANE)
from .ocean import Ocean, RESOLUTION, DOMAIN_SIZE  # noqa: F401
