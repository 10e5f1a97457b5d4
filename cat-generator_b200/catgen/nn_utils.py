"""Mirror of the sampler-side helpers of the reference's utils/nn_utils.lua over libcatgen.

Reference: /root/reference/utils/nn_utils.lua:35-39 (createNoiseInputs), :45-69 (createImagesFromNoise), :75-77 (createImages),
:89-117 (sortImagesByPrediction), :334-349 (switchToTrainingMode / switchToEvaluationMode); callers sample.lua:89-112 (G and D in
evaluate() mode, sample.lua:211,216) and adversarial.lua:221-238 (training mode).  Only host-side chunking and sorting live here;
every forward pass is libcatgen's CUDA path.  The reference reads OPT.batchSize / OPT.noiseDim / MODEL_G / MODEL_D from globals;
here they are arguments.

Differences that are deliberate and visible:
  * noise comes from a numpy Generator, not Torch's MT19937 stream (the reference's exact draws cannot be reproduced without Torch);
  * Lua's table.sort is not stable, so the reference's order among EQUAL predictions is unspecified; this sort is stable.
"""
import numpy as np

from .lib import f32


def createNoiseInputs(N, noiseDim=100, rng=None):
    """nn_utils.lua:35-39: [N, noiseDim] float tensor, uniform in [-1, 1)."""
    rng = np.random.default_rng() if rng is None else rng
    return rng.uniform(-1.0, 1.0, (int(N), int(noiseDim))).astype(np.float32)


def createImagesFromNoise(MODEL_G, noiseInputs, batchSize, outputAsList=False):
    """nn_utils.lua:45-69: G forward in chunks of batchSize (the last chunk may be smaller), each chunk copied into one [N,C,H,W]
    float tensor.  G's mode is whatever the caller left it in: batch-statistics BN (and a running-statistics update per chunk) in
    training mode, running statistics after evaluate()."""
    z = f32(noiseInputs)
    N = z.shape[0]
    images = None
    for start in range(0, N, int(batchSize)):
        end = min(start + int(batchSize), N)
        generated = MODEL_G.forward(z[start:end])
        if images is None:
            images = np.empty((N,) + generated.shape[1:], np.float32)
        images[start:end] = generated
    if images is None:
        images = np.empty((0, MODEL_G.C, 32, 32), np.float32)
    return [images[i] for i in range(N)] if outputAsList else images


def createImages(MODEL_G, N, batchSize, noiseDim=100, outputAsList=False, rng=None):
    """nn_utils.lua:75-77."""
    return createImagesFromNoise(MODEL_G, createNoiseInputs(N, noiseDim, rng), batchSize, outputAsList)


def sortImagesByPrediction(MODEL_D, images, ascending, nbMaxOut, batchSize):
    """nn_utils.lua:89-117: D's prediction per image (1.0 = "probably real"), chunked by batchSize; images sorted by it,
    descending unless `ascending`; at most nbMaxOut returned as (list of images, list of predictions)."""
    images = f32(images)
    N = images.shape[0]
    predictions = np.empty(N, np.float32)
    for start in range(0, N, int(batchSize)):
        end = min(start + int(batchSize), N)
        predictions[start:end] = MODEL_D.forward(images[start:end])[:, 0]
    order = np.argsort(predictions if ascending else -predictions, kind="stable")
    keep = order[:min(int(nbMaxOut), N)]
    return [images[i] for i in keep], [float(predictions[i]) for i in keep]


def switchToTrainingMode(MODEL_G, MODEL_D):
    """nn_utils.lua:334-340 (the auto-encoder branch is outside the hot path)."""
    MODEL_G.training(); MODEL_D.training()


def switchToEvaluationMode(MODEL_G, MODEL_D):
    """nn_utils.lua:343-349."""
    MODEL_G.evaluate(); MODEL_D.evaluate()


def rateWithV(MODEL_V, images):
    """NN_UTILS.rateWithV (utils/nn_utils.lua:686-711): 1 - mean P(fake) under V; accepts a list of images or one array."""
    x = np.stack([np.asarray(i, np.float32) for i in images]) if isinstance(images, (list, tuple)) else np.asarray(images, np.float32)
    predictions = MODEL_V.forward(x)
    return 1.0 - float(predictions[:, 0].astype(np.float64).mean())
