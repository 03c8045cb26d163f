from .sampler import Sampler  # noqa: F401
