"""Inference throughput of the sampler path (sample.lua:89-112): NN_UTILS.createImages(1024) with G in
evaluate() mode, then NN_UTILS.sortImagesByPrediction with D in evaluate() mode, chunks of OPT.batchSize = 128, through the
host-pointer C-ABI (H2D of the noise, D2H of the images, H2D of the images again for D, D2H of the predictions: as the
reference's Copy layers do).  Wall clock around synchronous calls; not a bench.py line.
    python tools/sampler_bench.py [N batchSize]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cat-generator_b200")]
from catgen import lib, models, nn_utils

def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    bs = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    lib.init(0)
    g = models.create_G((3, 32, 32), 100); d = models.create_D((3, 32, 32), True)
    rng = np.random.default_rng(0)
    g.forward(nn_utils.createNoiseInputs(bs, 100, rng))          # one training-mode forward: running statistics off their initial values
    nn_utils.switchToEvaluationMode(g, d)
    z = nn_utils.createNoiseInputs(N, 100, rng)
    for _ in range(2):
        images = nn_utils.createImagesFromNoise(g, z, bs); nn_utils.sortImagesByPrediction(d, images, False, 64, bs)
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps): images = nn_utils.createImagesFromNoise(g, z, bs)
    t1 = time.perf_counter()
    for _ in range(reps): best, preds = nn_utils.sortImagesByPrediction(d, images, False, 64, bs)
    t2 = time.perf_counter()
    print("sampler path, N=%d, batch %d, RGB 32x32, eval mode, host buffers in/out:" % (N, bs))
    print("  createImages            %8.1f images/s  (%.2f ms per %d)" % (reps * N / (t1 - t0), 1e3 * (t1 - t0) / reps, N))
    print("  sortImagesByPrediction  %8.1f images/s  (%.2f ms per %d)" % (reps * N / (t2 - t1), 1e3 * (t2 - t1) / reps, N))
    print("  both                    %8.1f images/s" % (reps * N / (t2 - t0)))
    print("  best prediction %.4f, 64th %.4f" % (preds[0], preds[-1]))

if __name__ == "__main__":
    main()
