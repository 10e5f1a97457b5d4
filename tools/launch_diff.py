"""Diagnosis: which launches differ between an eager step and a graph-captured step at a small batch (run with CATGEN_LAUNCH_TRACE=1).
    CATGEN_LAUNCH_TRACE=1 python tools/launch_diff.py 2> trace.txt; python tools/launch_diff.py --diff trace.txt"""
import collections, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cat-generator_b200"), os.path.join(ROOT, "tests")]

def run():
    import numpy as np
    from catgen import lib, models, adversarial
    L = lib.load(); lib.init(0)
    B, Cc = 8, 3
    rng = np.random.default_rng(11)
    for mode in (0, 1):
        lib.check(L.cg_set_graph_mode(mode))
        g = models.create_G((Cc, 32, 32), 100, seed=1); d = models.create_D((Cc, 32, 32), True, seed=2)
        t = adversarial.Trainer(g, d)
        for i in range(5):
            real = rng.uniform(0, 1, (1, B // 2, Cc, 32, 32)).astype(np.float32); zD = rng.uniform(-1, 1, (1, B // 2, 100)).astype(np.float32); zG = rng.uniform(-1, 1, (1, B, 100)).astype(np.float32)
            n0 = L.cg_launch_count()
            sys.stderr.write("[mark] mode %d step %d begin\n" % (mode, i)); sys.stderr.flush()
            t.step(lib.default_cfg(B), real, zD, zG)
            sys.stderr.write("[mark] mode %d step %d end count %d\n" % (mode, i, L.cg_launch_count() - n0)); sys.stderr.flush()

def diff(path):
    steps, cur = {}, None
    for line in open(path):
        if line.startswith("[mark]") and "begin" in line:
            p = line.split(); cur = (int(p[2]), int(p[4])); steps[cur] = []
        elif line.startswith("[mark]") and "end" in line:
            print(line.strip()); cur = None
        elif line.startswith("[launch]") and cur is not None:
            steps[cur].append(line.split()[1])
    ref = collections.Counter(steps[(0, 3)])
    for k in sorted(steps):
        c = collections.Counter(steps[k])
        d = {n: c[n] - ref[n] for n in set(c) | set(ref) if c[n] != ref[n]}
        print(k, len(steps[k]), "named launches; vs eager step 3:", d)

if __name__ == "__main__":
    diff(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[1] == "--diff" else run()
