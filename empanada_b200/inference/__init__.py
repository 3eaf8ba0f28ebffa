from empanada_b200.inference import postprocess, engines, rle  # noqa: F401
