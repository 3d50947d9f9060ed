"""Command-line entry point, flag-compatible with the reference train.py
(``--config``, ``--run-id``, ``--cpu``).  The reference's own train.py also drives this package
unchanged (PYTHONPATH pointing here): it only needs ``trainer.PPOTrainer`` and ``yaml_parser.YamlParser``."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from trainer import PPOTrainer  # noqa: E402
from yaml_parser import YamlParser  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser(description="PPO + TransformerXL episodic memory, B200-native engine")
    default_cfg = os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs", "poc_memory_env.yaml")
    ap.add_argument("--config", default=default_cfg, help="path to the yaml config file")
    ap.add_argument("--run-id", default="run", help="tag for the tensorboard summary and the saved model")
    ap.add_argument("--cpu", action="store_true", help="accepted for compatibility; this engine has no CPU path and will refuse")
    args = ap.parse_args(argv)
    config = YamlParser(args.config).get_config()
    # one process per GPU under torchrun: rank r trains on cuda:LOCAL_RANK with its own workers; gradients are all-reduced
    import parallel
    parallel.init_from_env()
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.cpu or not torch.cuda.is_available():
        device = torch.device("cpu")
    else:
        torch.cuda.set_device(local_rank)
        device = torch.device("cuda", local_rank)
    trainer = PPOTrainer(config, run_id=args.run_id, device=device)      # summaries / the saved model are written by rank 0 only
    trainer.run_training()
    trainer.close()


if __name__ == "__main__":
    main()
