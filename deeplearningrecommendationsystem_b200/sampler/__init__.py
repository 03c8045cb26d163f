from .sampler import Sampler  # noqa: F401
from .device import DeviceSampler  # noqa: F401
