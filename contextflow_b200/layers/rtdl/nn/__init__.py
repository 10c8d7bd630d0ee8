from ._embeddings import OneHotEncoder, EyeEncoder, CatEmbeddings  # noqa: F401
