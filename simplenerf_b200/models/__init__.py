from .ModelFactory import get_model  # noqa: F401
