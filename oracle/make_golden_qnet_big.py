"""Large action-parity fixture for the Q-network: >= 4096 states per agent window drawn from real greedy
episodes, with the UNMODIFIED reference Net's (eval mode) greedy action, best Q-value and top-2 margin
per state -> tests/golden/qnet_states_big.npz.

TEST INFRASTRUCTURE (container only: imports the reference through oracle/ref_shim.py).
    python oracle/make_golden_qnet_big.py

Episodes start from windows of the kind GA.mutation extracts (GA.py:338-350): rows x cols cuts of boards
made by the GA.generate_population recipe (gaps inserted into related sequences), so the token statistics
are those of the mutation phase, not uniform noise. Every state along the reference's own greedy
trajectory (dqn.py:151-154, env.py:231-259) is recorded until the quota is full.
"""
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ga_oracle as O  # noqa: E402
import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
WEIGHT_SEED = 7
QUOTA = 4096


def related_sequences(rng, rows, cols):
    base = [rng.randint(1, 4) for _ in range(cols)]
    seqs = []
    for _ in range(rows):
        s = [b if rng.random() > 0.15 else rng.randint(1, 4) for b in base]
        seqs.append("".join("ATCG"[v - 1] for v in s))
    return seqs


def main():
    import torch
    ref = ref_shim.load()
    rng = random.Random(4242)
    pack = {}
    for rows, cols in ((3, 30), (6, 30)):
        sd = O.synth_qnet_weights(rows, cols, WEIGHT_SEED)
        max_value = cols * rows * (rows - 1) / 2 * 4
        agent = ref.dqn.DQN(2 ** rows - 1, rows, cols, max_value)
        agent.eval_net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        ref_shim.force_eval(agent)
        states = []
        seen = set()
        while len(states) < QUOTA:
            seqs = related_sequences(rng, rows, cols + 6)
            board = O.generate_population(seqs, 2, 0.0, 0.08, rng)[0]
            fc = rng.randint(0, min(len(r) for r in board) - cols)
            env = O.Env([row[fc:fc + cols] for row in board])
            s = env.reset()
            while True:
                key = bytes(s)
                if key not in seen:
                    seen.add(key)
                    states.append(list(s))
                with torch.no_grad():
                    a = int(agent.predict(s))
                _, s, d = env.step(a)
                if d == 0:
                    break
        states = np.asarray(states[:QUOTA], dtype=np.int8)
        with torch.no_grad():
            q = agent.eval_net(torch.from_numpy(states.astype(np.int64))).numpy().astype(np.float32)
        order = np.argsort(-q, axis=1, kind="stable")
        best = order[:, 0]
        top1 = q[np.arange(len(q)), best]
        top2 = q[np.arange(len(q)), order[:, 1]]
        assert np.array_equal(best, q.argmax(1))
        tag = "%dx%d" % (rows, cols)
        pack["states_" + tag] = states
        pack["action_" + tag] = best.astype(np.uint8)
        pack["top1_" + tag] = top1
        pack["margin_" + tag] = (top1 - top2).astype(np.float32)
        pack["q_head_" + tag] = q[:256]
        print(tag, states.shape, "min margin %.3g (%.3g of max_value), median %.3g" %
              (float((top1 - top2).min()), float((top1 - top2).min()) / max_value, float(np.median(top1 - top2))))
    np.savez_compressed(os.path.join(OUT, "qnet_states_big.npz"), weight_seed=WEIGHT_SEED, **pack)
    print("written", os.path.getsize(os.path.join(OUT, "qnet_states_big.npz")), "bytes")


if __name__ == "__main__":
    main()
