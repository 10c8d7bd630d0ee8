"""Namespace shim: model.py imports `layers.rtdl.nn._embeddings` (reference model.py:14).  Only the three categorical
embedding modules that ContextEncoder reaches are provided; the rest of the vendored rtdl library is outside the path."""
