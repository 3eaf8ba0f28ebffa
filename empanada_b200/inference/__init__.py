from empanada_b200.inference import postprocess, engines, rle, matcher  # noqa: F401
