"""GA utilities with the reference's names and positional signatures (reference utils.py:27-469),
computed by libgadpamsa on the GPU. Boards passed to these free functions are the reference's
python lists; they are copied to the device, scored there and the result is returned as the
same python type the reference returns (int SP, float CS, list of indices, ...).

There is no CPU fallback: without the CUDA library or a GPU these functions raise.
The callers' half of the reference's utils.py (utils.py:472-975) follows at the bottom: metrics, FASTA
parsing and the CSV plumbing are host formatting around device-computed scores; the external-tool runner is
plain subprocess plumbing (the tools themselves are not shipped) and plot_metrics is a reduced version.
"""
from __future__ import annotations

import ctypes

import numpy as np

import config
import _native as nat

_LETTERS = "ATCG-"          # reference utils.py:458


# ----------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------
def _pack_window(chromosome, from_row, to_row, from_column, to_column):
    """The half-open window of a list-of-lists board as a dense uint8 [rows, stride] array."""
    rows, cols = to_row - from_row, to_column - from_column
    stride = max(4, -(-cols // 4) * 4)
    arr = np.full((rows, stride), 5, dtype=np.uint8)
    for i in range(rows):
        row = chromosome[from_row + i]
        if to_column > len(row):
            raise IndexError("list index out of range")          # what the reference raises
        arr[i, :cols] = row[from_column:to_column]
    return arr, stride


def _window_score(chromosome, from_row, to_row, from_column, to_column):
    """(SP, exact-match columns) of a non-empty window, scored on the device."""
    nat.require_device()
    arr, stride = _pack_window(chromosome, from_row, to_row, from_column, to_column)
    sp = ctypes.c_int64(0)
    em = ctypes.c_int32(0)
    nat.call("gad_window_score_host", arr.ctypes.data, arr.shape[0], stride, 0, to_row - from_row, 0,
             to_column - from_column, int(config.MATCH_REWARD), int(config.MISMATCH_PENALTY),
             int(config.GAP_PENALTY), ctypes.byref(sp), ctypes.byref(em))
    return sp.value, em.value


# ----------------------------------------------------------------------------------------------
# a6 / a7: scoring (reference utils.py:27-128)
# ----------------------------------------------------------------------------------------------
def get_sum_of_pairs(chromosome, from_row, to_row, from_column, to_column):
    """Sum-of-pairs of a half-open window (reference utils.py:62-76). Returns a python int."""
    if to_row - from_row < 2 or to_column <= from_column:
        return 0
    return int(_window_score(chromosome, from_row, to_row, from_column, to_column)[0])


def get_column_score(chromosome, from_row, to_row, from_column, to_column):
    """Fraction of columns whose cells all equal the first row's (reference utils.py:119-128);
    python true division of device-counted integers, 0 when the window has no columns."""
    num_columns = to_column - from_column
    if num_columns <= 0:
        return 0
    if to_row <= from_row:
        return num_columns / num_columns          # all() over no rows is True for every column
    em = _window_score(chromosome, from_row, to_row, from_column, to_column)[1]
    return em / num_columns


# ----------------------------------------------------------------------------------------------
# a5: all-gap column removal (reference utils.py:390-428), in place on the caller's lists
# ----------------------------------------------------------------------------------------------
def check_if_there_are_all_gaps(row, from_index):
    """reference utils.py:383-387 (tiny host helper kept for API completeness)."""
    return False if any(v != 5 for v in row[from_index:]) else from_index


def clean_unnecessary_gaps(aligned_sequence):
    """Remove every all-gap column of a rectangular board, in place (reference utils.py:404-428)."""
    if not aligned_sequence or not aligned_sequence[0]:
        return
    nat.require_device()
    from population import pack_boards
    width = len(aligned_sequence[0])
    if any(len(r) != width for r in aligned_sequence):
        raise IndexError("list index out of range")              # reference behaviour on ragged input
    boards, rowlen = pack_boards([aligned_sequence])
    P, R, stride = boards.shape
    out = np.zeros(3, dtype=np.int32)
    nat.call("gad_fitness_host", boards.ctypes.data, rowlen.ctypes.data, 1, R, stride,
             int(config.MATCH_REWARD), int(config.MISMATCH_PENALTY), int(config.GAP_PENALTY),
             out[0:1].ctypes.data, out[1:2].ctypes.data, out[2:3].ctypes.data, 1)
    new_w = int(rowlen[0, 0])
    for r, row in enumerate(aligned_sequence):
        row[:] = boards[0, r, :new_w].tolist()


def get_nucleotides_seqs(chromosome):
    """int codes -> strings (reference utils.py:464-467); pure formatting, stays on the host."""
    return ["".join(_LETTERS[v - 1] for v in sequence) for sequence in chromosome]


def count_exact_columns(chromosome, from_row, to_row, from_column, to_column):
    """Number of exact-match columns of a window (numerator of utils.py:128), device-counted."""
    if to_column <= from_column:
        return 0
    if to_row <= from_row:
        return to_column - from_column
    return int(_window_score(chromosome, from_row, to_row, from_column, to_column)[1])


# ----------------------------------------------------------------------------------------------
# a15: tiling and the worst-fitted sub-board (reference utils.py:131-313)
# ----------------------------------------------------------------------------------------------
def is_overlap(range1, range2):
    """reference utils.py:153-162 (half-open rectangles)."""
    from_row1, to_row1, from_column1, to_column1 = range1
    from_row2, to_row2, from_column2, to_column2 = range2
    return (from_row1 < to_row2 and to_row1 > from_row2) and (from_column1 < to_column2 and to_column1 > from_column2)


def check_overlap(new_range, used_ranges):
    """reference utils.py:191-195"""
    return any(is_overlap(new_range, r) for r in used_ranges)


def get_all_different_sub_range(individual, m_prime=None, n_prime=None):
    """The non-overlapping full tiles the reference's scan accepts (utils.py:230-252): exactly the
    lattice of (m_prime, n_prime) tiles that fit, row-major. Defaults are read from config."""
    m_prime = config.AGENT_WINDOW_ROW if m_prime is None else m_prime
    n_prime = config.AGENT_WINDOW_COLUMN if n_prime is None else n_prime
    m = len(individual)
    n = min(len(genes) for genes in individual)
    return [(i, i + m_prime, j, j + n_prime)
            for i in range(0, m - m_prime + 1, m_prime)
            for j in range(0, n - n_prime + 1, n_prime)]


def calculate_worst_fitted_sub_board(individual, mode):
    """(score, (from_row, to_row, from_column, to_column)) of the lowest-scoring tile (reference
    utils.py:276-313). The tile is chosen by gad_worst_tiles on the device; the returned score is the
    reference's first tuple element: the tile's SP ('sp' and 'cs' modes) or its normalised SP ('mo')."""
    import torch
    from population import DevicePopulation
    rw, cw = int(config.AGENT_WINDOW_ROW), int(config.AGENT_WINDOW_COLUMN)
    tiles = get_all_different_sub_range(individual, rw, cw)
    if not tiles:
        raise ValueError("min() arg is an empty sequence")
    dev = DevicePopulation.from_lists([individual], device=config.DEVICE)
    t = dev.worst_tiles(None, rw, cw, mode, int(config.MATCH_REWARD), int(config.MISMATCH_PENALTY),
                        int(config.GAP_PENALTY)).cpu().numpy()[0]
    torch.cuda.synchronize()
    fr, fc = int(t[0]), int(t[1])
    worst = (fr, fr + rw, fc, fc + cw)
    sp = get_sum_of_pairs(individual, *worst)
    if mode in ('sp', 'cs'):
        return sp, worst
    sps = [get_sum_of_pairs(individual, *tile) for tile in tiles]
    lo, hi = min(sps), max(sps)
    return ((sp - lo) / (hi - lo) if hi > lo else 0.5), worst


# ----------------------------------------------------------------------------------------------
# a9: ranking (reference utils.py:332-355)
# ----------------------------------------------------------------------------------------------
def get_index_of_the_best_fitted_individuals(population_scores, num_individuals):
    """Indices of the best individuals: stable descending sort on the score ('sp'/'cs' tuples) or on
    the sum of the min/max-normalised SP and CS ('mo' tuples). The keys are formed on the host with the
    reference's own float expressions; the stable sort runs on the device (gad_rank_order)."""
    import torch
    n = len(population_scores)
    if n == 0:
        return []
    nat.require_device()
    if len(population_scores[0]) == 2:
        keys = np.array([float(t[1]) for t in population_scores], dtype=np.float64)
    else:
        sps = [t[1] for t in population_scores]
        css = [t[2] for t in population_scores]
        lo_s, hi_s, lo_c, hi_c = min(sps), max(sps), min(css), max(css)
        keys = np.array([((s - lo_s) / (hi_s - lo_s) if hi_s > lo_s else 0.5) +
                         ((c - lo_c) / (hi_c - lo_c) if hi_c > lo_c else 0.5)
                         for s, c in zip(sps, css)], dtype=np.float64)
    dev = torch.device(config.DEVICE)
    nbytes = ctypes.c_size_t(0)
    nat.call("gad_workspace_bytes", n, ctypes.byref(nbytes))
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
    d_keys = torch.from_numpy(keys).to(dev)
    d_order = torch.empty(n, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        nat.call("gad_rank_order", d_keys.data_ptr(), n, d_order.data_ptr(), ws.data_ptr(), ws.numel(),
                 nat.current_stream_ptr())
    order = d_order.cpu().numpy()
    return [population_scores[i][0] for i in order[:num_individuals]]


# ----------------------------------------------------------------------------------------------
# callers' side of the path: metrics, FASTA parsing, report/CSV plumbing (reference utils.py:475-807).
# Host-side formatting only; the scores themselves come from the device functions above.
# ----------------------------------------------------------------------------------------------
def calculate_metrics(env):
    """reference utils.py:500-513"""
    alignment_length = len(env.aligned[0])
    num_sequences = len(env.aligned)
    sp_score = env.calc_score()
    exact_matches = env.calc_exact_matched()
    column_score = exact_matches / alignment_length
    return {"AL": alignment_length, "QTY": num_sequences, "SP": sp_score, "EM": exact_matches, "CS": column_score}


def parse_fasta_to_sequences(fasta_content):
    """reference utils.py:543-560"""
    sequences, current = [], []
    for line in fasta_content.splitlines():
        line = line.strip()
        if line.startswith(">"):
            if current:
                sequences.append(''.join(current))
                current = []
        else:
            current.append(line)
    if current:
        sequences.append(''.join(current))
    return sequences


def save_inference_csv(csv_data, tool_name, dataset_name):
    """reference utils.py:733-748"""
    import os
    import pandas as pd
    tool_csv_dir = os.path.join(config.CSV_PATH, tool_name)
    os.makedirs(tool_csv_dir, exist_ok=True)
    csv_file_path = os.path.join(tool_csv_dir, f"{dataset_name}_{tool_name}_results.csv")
    if isinstance(csv_data, list):
        columns = ["File Name", "Alignment Length (AL)", "Number of Sequences (QTY)", "Sum of Pairs (SP)",
                   "Exact Matches (EM)", "Column Score (CS)"]
        df = pd.DataFrame(csv_data, columns=columns)
    else:
        df = pd.read_csv(csv_data)
    df.to_csv(csv_file_path, index=False)
    return csv_file_path


def run_ga_dpamsa_inference(mode, dataset, dataset_name, model_path):
    """reference utils.py:773-780"""
    import os
    from mainGA import inference as ga_inference
    ga_inference(mode=mode, dataset=dataset, model_path=model_path)
    mode_tag = {"sp": "Max_SP", "cs": "Max_CS", "mo": "MO"}[mode]
    return os.path.join(config.GA_DPAMSA_INF_CSV_PATH, f"{dataset_name}_{mode_tag}_GA_DPAMSA_results.csv")


CSV_COLUMNS_QTY_FIRST = ["File Name", "Number of Sequences (QTY)", "Alignment Length (AL)", "Sum of Pairs (SP)",
                         "Exact Matches (EM)", "Column Score (CS)"]


def report_block(name, metrics, alignment):
    """One record of the reference's report files (mainGA.py:137-145, DPAMSA/main.py:291-299, utils.py:677-683) -
    the format of the 1 340 published records."""
    return (f"File: {name}\n"
            f"Number of Sequences (QTY): {metrics['QTY']}\n"
            f"Alignment Length (AL): {metrics['AL']}\n"
            f"Sum of Pairs (SP): {metrics['SP']}\n"
            f"Exact Matches (EM): {metrics['EM']}\n"
            f"Column Score (CS): {metrics['CS']:.3f}\n"
            f"Alignment:\n{alignment}\n\n")


def run_dpamsa_inference(dataset, dataset_name, model_path):
    """reference utils.py:801-807: plain DPAMSA (no GA) over a dataset module; returns the CSV path."""
    import os
    from DPAMSA.main import inference as dpamsa_inference
    dpamsa_inference(dataset=dataset, model_path=model_path, truncate_file=True)
    return os.path.join(config.DPAMSA_INF_CSV_PATH, f"{dataset_name}_DPAMSA_results.csv")


def display_menu():
    """reference utils.py:593-607: the three benchmarking options, asked on stdin until the answer is valid."""
    options = {1: "GA-DPAMSA vs DPAMSA", 2: "GA-DPAMSA vs Other MSA Tools", 3: "GA-DPAMSA vs DPAMSA vs Other MSA Tools"}
    print("Select the benchmarking option:")
    for key, text in options.items():
        print(f"{key}. {text}")
    while True:
        answer = input("Enter your choice (1, 2, or 3): ")
        try:
            choice = int(answer)
        except ValueError:
            print("Invalid input. Please enter a number.")
            continue
        if choice in options:
            return choice
        print("Please select a valid option (1, 2, or 3).")


def run_tool_and_generate_report(tool_name, file_paths, dataset_name):
    """reference utils.py:629-697: run an external MSA tool (config.TOOLS[tool_name]) on FASTA files, score its
    alignments on the GPU and write the report; returns [file, QTY, AL, SP, EM, CS] rows. Host plumbing around
    subprocess - the tools themselves are not part of this build (a missing binary raises FileNotFoundError)."""
    import glob
    import os
    import subprocess
    from DPAMSA.env import Environment
    tool = config.TOOLS[tool_name]
    out_dir = os.path.join(tool['output_dir'], dataset_name)
    os.makedirs(out_dir, exist_ok=True)
    os.makedirs(tool['report_dir'], exist_ok=True)
    rows = []
    with open(os.path.join(tool['report_dir'], f"{dataset_name}.txt"), 'w') as report:
        for path in file_paths:
            fname = os.path.basename(path)
            stem = os.path.splitext(fname)[0]
            command = tool['command'](path, os.path.join(out_dir, fname))
            if isinstance(command, str):                       # shell pipeline (MAFFT in the reference)
                done = subprocess.run(command, shell=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            else:
                done = subprocess.run(command, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            produced = {"UPP": os.path.join(out_dir, fname, "output_alignment.fasta"),
                        "PASTA": os.path.join(out_dir, fname, f"pastajob.marker001.{stem}.aln")}.get(
                            tool_name, os.path.join(out_dir, fname))
            if os.path.exists(produced):
                with open(produced) as fh:
                    fasta = fh.read()
            else:
                fasta = done.stdout or ""
            aligned = parse_fasta_to_sequences(fasta)
            env = Environment(aligned, convert_data=False)
            Environment.set_alignment(env, aligned)
            m = calculate_metrics(env)
            report.write(report_block(fname, m, env.get_alignment()))
            rows.append([fname, m['QTY'], m['AL'], m['SP'], m['EM'], m['CS']])
            if tool_name == 'ClustalW':
                for dnd in glob.glob(os.path.join(os.path.dirname(path), '*.dnd')):
                    os.remove(dnd)
    return rows


def plot_metrics(tool_csv_paths, dataset_name):
    """reference utils.py:813-975 draws box/bar charts with matplotlib. Plotting is outside the scope of this
    build: with matplotlib installed a compact SP / CS bar chart per tool is written under config.CHARTS_PATH,
    without it the per-tool means are printed. Returns {tool: {"SP": mean, "CS": mean}}."""
    import os
    import pandas as pd
    summary = {}
    for tool, path in tool_csv_paths.items():
        df = pd.read_csv(path)
        summary[tool] = {"SP": float(df["Sum of Pairs (SP)"].mean()), "CS": float(df["Column Score (CS)"].mean())}
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except ImportError:
        for tool, m in summary.items():
            print(f"{dataset_name} {tool}: mean SP {m['SP']:.2f}, mean CS {m['CS']:.4f}")
        return summary
    os.makedirs(config.CHARTS_PATH, exist_ok=True)
    for metric in ("SP", "CS"):
        fig, ax = plt.subplots(figsize=(max(4, len(summary)), 3))
        ax.bar(list(summary), [m[metric] for m in summary.values()])
        ax.set_title(f"{dataset_name}: mean {metric}")
        fig.tight_layout()
        fig.savefig(os.path.join(config.CHARTS_PATH, f"{dataset_name}_{metric}.png"))
        plt.close(fig)
    return summary
