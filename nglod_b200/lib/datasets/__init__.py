from .MeshDataset import MeshDataset  # noqa: F401
from .SPCDataset import SPCDataset  # noqa: F401
