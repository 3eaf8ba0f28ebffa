from empanada_b200.inference import postprocess, engines, rle, matcher, fill  # noqa: F401
