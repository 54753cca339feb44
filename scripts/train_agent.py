"""Train a DPAMSA agent with this package's trainer (DPAMSA.main.train, reference DPAMSA/main.py:95-235) on a
dataset module in the reference's format (a .py file with ``file_name``, ``datasets`` and one list of sequences per
name) and save ``<model>.pth`` under config.DPAMSA_WEIGHTS_PATH, ready for ``mainGA.inference`` / ``GA.run``.

    python scripts/train_agent.py path/to/zhang_dataset_3x30.py --model new_model_3x30 --episodes 6000 [--start 0 --end -1]

Needs a GPU (the environment's rewards come from the device scorer). Knobs not given here are config.py's.
"""
import argparse
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "ga-dpamsa_b200"), ROOT]


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("dataset", help="dataset module file (reference datasets/training_dataset/*.py format)")
    ap.add_argument("--model", default="new_model_3x30")
    ap.add_argument("--episodes", type=int, default=None, help="config.MAX_EPISODE")
    ap.add_argument("--batch-size", type=int, default=None)
    ap.add_argument("--replay-size", type=int, default=None)
    ap.add_argument("--lr", type=float, default=None, help="config.ALPHA")
    ap.add_argument("--start", type=int, default=0)
    ap.add_argument("--end", type=int, default=-1)
    ap.add_argument("--patience", type=int, default=200, help="early stopping: episodes without a better 100-episode mean")
    args = ap.parse_args()

    import math
    import config
    if args.episodes is not None:
        config.MAX_EPISODE = args.episodes
        config.DECREMENT_ITERATION = math.ceil(config.MAX_EPISODE * 0.8 / (config.EPSILON // config.DELTA))
    if args.batch_size is not None:
        config.BATCH_SIZE = args.batch_size
    if args.replay_size is not None:
        config.REPLAY_MEMORY_SIZE = args.replay_size
    if args.lr is not None:
        config.ALPHA = args.lr
    assert 0 < config.BATCH_SIZE <= config.REPLAY_MEMORY_SIZE, "batch size must be in (0, replay memory size]"

    spec = importlib.util.spec_from_file_location("_train_data", args.dataset)
    data = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(data)
    if not hasattr(data, "file_name"):
        data.file_name = os.path.basename(args.dataset)
    if not hasattr(data, "datasets"):
        data.datasets = [k for k, v in vars(data).items() if isinstance(v, list) and v and isinstance(v[0], str)]

    import DPAMSA.main as dmain
    os.makedirs(config.DPAMSA_WEIGHTS_PATH, exist_ok=True)
    history = dmain.train(data, start=args.start, end=args.end, model_path=args.model,
                          early_stopping_patience=args.patience)
    for name, log in history.items():
        tail = log[-100:]
        print("%s: %d episodes, mean reward of the last %d: %.2f, last SP %s" %
              (name, len(log), len(tail), sum(e["reward"] for e in tail) / max(1, len(tail)), log[-1]["sp"] if log else "-"))


if __name__ == "__main__":
    main()
