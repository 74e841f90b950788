from .OctreeSDF import OctreeSDF, FeatureVolume  # noqa: F401
from .BaseLOD import BaseLOD  # noqa: F401
from .BaseSDF import BaseSDF  # noqa: F401
