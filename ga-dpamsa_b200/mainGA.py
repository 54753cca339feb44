"""GA-DPAMSA inference driver with the reference script's interface (reference mainGA.py:70-153):
``inference(mode, dataset, start, end, model_path, debug, truncate_file)`` walks the sequence sets of a
dataset module, runs ``GA.run`` on the GPU for each, scores the winning alignment with the device
metrics (``utils.calculate_metrics``) and appends one block per set to the report and one row to the
CSV, in the formats the reference writes (and its ``plot_metrics`` / golden reports read).

The reference's own ``mainGA.py`` also works unchanged on top of this package; this
module exists so that ``utils.run_ga_dpamsa_inference`` and users without the reference checkout have
the same entry point. A dataset module is any object with ``file_name``, ``datasets`` (names) and one
attribute per name holding the list of sequences (reference datasets/inference_dataset/*.py).
"""
from __future__ import annotations

import csv
import os

import config
import utils
from DPAMSA.env import Environment
from GA import GA

GA_MODE = 'sp'
DATASET = None
INFERENCE_MODEL = 'new_model_3x30'
DEBUG_MODE = False

MODE_TAGS = {"sp": "Max_SP", "cs": "Max_CS", "mo": "MO"}
CSV_HEADER = ["File Name", "Number of Sequences (QTY)", "Alignment Length (AL)", "Sum of Pairs (SP)",
              "Exact Matches (EM)", "Column Score (CS)"]


def output_parameters():
    print("\n-------- Genetic Algorithm parameters ---------")
    print(f"Window size:{config.AGENT_WINDOW_ROW}x{config.AGENT_WINDOW_COLUMN}")
    print(f"Population Size: {config.POPULATION_SIZE}")
    print(f"Number of iteration: {config.GA_ITERATIONS}")
    print(f"Clone Rate: {config.CLONE_RATE * 100}%")
    print(f"Gap Rate: {config.GAP_RATE * 100}%")
    print(f"Selection Rate: {config.SELECTION_RATE * 100}%")
    print(f"Mutation Rate: {config.MUTATION_RATE * 100}%\n")


def inference(mode, dataset=None, start=0, end=-1, model_path='model_3x30', debug=False, truncate_file=True):
    dataset = DATASET if dataset is None else dataset
    if mode not in MODE_TAGS:
        raise ValueError(f"Invalid mode '{mode}'. Choose one of {set(MODE_TAGS)}.")
    if dataset is None:
        raise ValueError("no dataset module given")
    tag = os.path.splitext(dataset.file_name)[0]
    os.makedirs(config.GA_DPAMSA_REPORTS_PATH, exist_ok=True)
    os.makedirs(config.GA_DPAMSA_INF_CSV_PATH, exist_ok=True)
    report_file_name = os.path.join(config.GA_DPAMSA_REPORTS_PATH, f"{tag}_{MODE_TAGS[mode]}.txt")
    csv_file_name = os.path.join(config.GA_DPAMSA_INF_CSV_PATH, f"{tag}_{MODE_TAGS[mode]}_GA_DPAMSA_results.csv")
    if truncate_file:
        open(report_file_name, 'w').close()
        with open(csv_file_name, 'w', newline='') as fh:
            csv.writer(fh).writerow(CSV_HEADER)
    output_parameters()
    names = dataset.datasets[start:end if end != -1 else len(dataset.datasets)]
    for name in names:
        if not hasattr(dataset, name):
            continue
        seqs = getattr(dataset, name)
        env = Environment(seqs, convert_data=False)
        best_alignment = GA(seqs, mode).run(model_path, debug)
        Environment.set_alignment(env, best_alignment)
        m = utils.calculate_metrics(env)
        with open(report_file_name, 'a') as fh:
            fh.write(utils.report_block(name, m, env.get_alignment()))
        with open(csv_file_name, 'a', newline='') as fh:
            csv.writer(fh).writerow([name, m['QTY'], m['AL'], m['SP'], m['EM'], m['CS']])
    print(f"\nInference completed successfully.\nReport saved at: {report_file_name}\nCSV saved at: {csv_file_name}\n")
    return report_file_name, csv_file_name


if __name__ == "__main__":
    import importlib
    import sys
    module = sys.argv[1] if len(sys.argv) > 1 else "datasets.inference_dataset.dataset1_3x30bp"
    inference(mode=GA_MODE, dataset=importlib.import_module(module), model_path=INFERENCE_MODEL, debug=DEBUG_MODE)
