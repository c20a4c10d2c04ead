"""Command line of the zonal pipeline: drop-in for flair_zonal_detection/main.py (``--config <zonal config>``).

    python -m flair_for_aigle_b200.flair_zonal_detection.main --config configs/config_model_zonal_segmentation.yaml
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 -m flair_for_aigle_b200.flair_zonal_detection.main --config ...

Under torchrun (WORLD_SIZE > 1) the process group is created here, one process per GPU, and ``run_inference`` splits the zone
into row strips over the ranks (the reference runs on one device); rank 0 writes the output rasters.
"""
from __future__ import annotations

import argparse
import logging
import os


def main(argv=None) -> None:
    parser = argparse.ArgumentParser(description="Run zonal detection inference.")
    parser.add_argument("--config", type=str, required=True, help="Path to the detection config file")
    args = parser.parse_args(argv)
    logging.basicConfig(level=logging.INFO, format="%(message)s")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        local = int(os.environ.get("LOCAL_RANK", "0"))
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group("gloo")
    try:
        from .inference import run_inference
        run_inference(args.config)
    finally:
        if dist is not None and dist.is_initialized():
            dist.barrier()
            dist.destroy_process_group()


if __name__ == '__main__':
    main()
