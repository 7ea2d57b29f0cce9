from .ms_deform_attn_func import MSDeformAttnFunction, MSDeformAttnFusedFunction, fused_supported  # noqa: F401
