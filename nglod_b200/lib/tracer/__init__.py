from .SphereTracer import SphereTracer  # noqa: F401
from .RenderBuffer import RenderBuffer  # noqa: F401
from .BaseTracer import BaseTracer  # noqa: F401
