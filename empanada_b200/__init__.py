"""empanada_b200 — B200 (sm_100a) implementation of empanada's panoptic inference post-processing
and 3D stack path behind the reference's own postprocess / engines / rle API.

    from empanada_b200.inference import postprocess, engines, rle

The CUDA library (empanada_b200/lib/libempanada_b200.so, built by ``python -m empanada_b200.build``)
is required: there is no CPU fallback.
"""
__version__ = '0.1.0'
