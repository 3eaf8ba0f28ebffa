from empanada_b200.inference import postprocess, engines, rle, matcher, fill, tracker, filters  # noqa: F401  (patterns imports consensus, which imports this package: import it by name)
