from .ms_deform_attn_func import (MSDeformAttnFunction, MSDeformAttnFusedFunction, ValueToBF16Function,  # noqa: F401
                                   bf16_value_mode, fused_supported)
