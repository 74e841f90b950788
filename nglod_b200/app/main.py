"""Train an OctreeSDF on a mesh -- the reference's app/main.py:44-48.

    python -m nglod_b200.app.main --net OctreeSDF --num-lods 5 --dataset-path mesh.obj --epochs 250 --exp-name run
    torchrun --nproc-per-node 8 -m nglod_b200.app.main ...      # data-parallel over the batch
"""
import logging as log

from ..lib.options import parse_options
from ..lib.trainer import Trainer

if __name__ == "__main__":
    log.basicConfig(format="%(asctime)s|%(levelname)8s| %(message)s", level=log.INFO)
    args, args_str = parse_options()
    log.info(f"Parameters: \n{args_str}")
    Trainer(args, args_str).train()
