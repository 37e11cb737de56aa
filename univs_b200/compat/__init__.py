"""Drop-in modules for a MinghanLi/UniVS checkout: put this directory on PYTHONPATH and the reference's
own `import MultiScaleDeformableAttention as MSDA` (ops/functions/ms_deform_attn_func.py:21) resolves to the B200 operator."""
