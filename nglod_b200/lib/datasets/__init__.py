from .MeshDataset import MeshDataset  # noqa: F401
