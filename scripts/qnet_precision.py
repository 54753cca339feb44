"""Q-value error of the two l1 implementations against the fp64 oracle (run on the GPU box).
Test infrastructure: imports oracle/ as the checker."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "ga-dpamsa_b200"), os.path.join(ROOT, "oracle")]
os.environ.setdefault("GADPAMSA_PROJECT_ROOT", "/tmp/gadpamsa_project_root")
import numpy as np
import ga_oracle as O
from DPAMSA.dqn import QNet

z = np.load(os.path.join(ROOT, "tests", "golden", "qnet_vectors.npz"))
for rows, cols in ((3, 30), (6, 30)):
    sd = O.synth_qnet_weights(rows, cols, int(z["weight_seed"]))
    states = z["states_%dx%d" % (rows, cols)].astype(np.uint8)
    mv = cols * rows * (rows - 1) / 2 * 4
    q64 = O.qnet_forward(sd, states, mv)
    qref = z["q_%dx%d" % (rows, cols)]
    out = {}
    for tag, env in (("simt", "1"), ("tc", "0")):
        os.environ["GAD_QNET_SIMT"] = env
        q, a = QNet(sd, rows, cols).forward(states, mv)
        out[tag] = q
        rel = np.abs(q - q64) / np.maximum(np.abs(q64), 1e-3 * mv)
        print("%dx%d %-5s vs fp64: max rel %.3e  mean rel %.3e  mean signed %.3e | argmax==ref %s" % (
            rows, cols, tag, rel.max(), rel.mean(), ((q - q64) / np.maximum(np.abs(q64), 1e-3 * mv)).mean(),
            np.array_equal(a, qref.argmax(1))))
    rel = np.abs(qref - q64) / np.maximum(np.abs(q64), 1e-3 * mv)
    print("%dx%d torch fp32 reference vs fp64: max rel %.3e mean rel %.3e" % (rows, cols, rel.max(), rel.mean()))
    srt = np.sort(q64, axis=1)
    print("   top-2 margin / |Q|: min %.3e median %.3e" % (((srt[:, -1] - srt[:, -2]) / np.abs(srt[:, -1])).min(),
                                                            np.median((srt[:, -1] - srt[:, -2]) / np.abs(srt[:, -1]))))
