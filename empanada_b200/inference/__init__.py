from empanada_b200.inference import postprocess, engines, rle, matcher, fill, tracker  # noqa: F401
