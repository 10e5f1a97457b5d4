"""Host-side logic that needs no GPU: the sampler helpers' chunking / sorting (catgen/nn_utils.py, mirroring
/root/reference/utils/nn_utils.lua:45-117) and the flat checkpoint (catgen/checkpoint.py), driven with stand-in networks."""
import numpy as np
import pytest

from catgen import checkpoint, nn_utils


class FakeG:
    kind, C, nz, nparams = 1, 3, 100, 7

    def __init__(self):
        self.calls, self.p, self.run, self.mode = [], np.arange(7, dtype=np.float32), np.ones(4, np.float32), "train"

    def forward(self, z):
        self.calls.append(z.shape[0])
        return np.broadcast_to(z[:, :1, None, None], (z.shape[0], 3, 32, 32)).astype(np.float32)   # image = first noise component

    def get_params(self): return self.p.copy()
    def set_params(self, a): self.p = np.asarray(a, np.float32).copy()
    def get_bn_running(self): return self.run.copy()
    def set_bn_running(self, a): self.run = np.asarray(a, np.float32).copy()
    def training(self): self.mode = "train"
    def evaluate(self): self.mode = "eval"


class FakeD(FakeG):
    nparams = 5

    def __init__(self):
        super().__init__(); self.p = np.arange(5, dtype=np.float32)

    def forward(self, x):
        self.calls.append(x.shape[0])
        return x[:, 0, 0, :1].astype(np.float32)      # prediction = the image's first pixel


def test_create_images_chunks_with_ragged_tail():
    g = FakeG()
    z = nn_utils.createNoiseInputs(21, 100, np.random.default_rng(0))
    assert z.shape == (21, 100) and z.dtype == np.float32 and -1 <= z.min() and z.max() < 1
    img = nn_utils.createImagesFromNoise(g, z, 8)
    assert g.calls == [8, 8, 5] and img.shape == (21, 3, 32, 32)           # ceil(N / batchSize) forwards, last one smaller
    assert np.array_equal(img[:, 0, 0, 0], z[:, 0])                         # chunks land at their own rows
    lst = nn_utils.createImagesFromNoise(g, z, 32, outputAsList=True)
    assert len(lst) == 21 and lst[3].shape == (3, 32, 32)
    assert nn_utils.createImagesFromNoise(g, z[:0], 8).shape == (0, 3, 32, 32)
    assert nn_utils.createImages(g, 5, 4, rng=np.random.default_rng(1)).shape == (5, 3, 32, 32)


def test_sort_images_by_prediction():
    d = FakeD()
    vals = np.array([0.2, 0.9, 0.5, 0.9, 0.1, 0.7, 0.3], np.float32)
    images = np.broadcast_to(vals[:, None, None, None], (7, 3, 32, 32)).astype(np.float32)
    best, p = nn_utils.sortImagesByPrediction(d, images, False, 3, 4)
    assert d.calls == [4, 3]
    assert p == pytest.approx([0.9, 0.9, 0.7]) and best[0][0, 0, 0] == np.float32(0.9)
    worst, q = nn_utils.sortImagesByPrediction(d, images, True, 64, 4)      # nbMaxOut larger than N: all of them
    assert len(worst) == 7 and q == pytest.approx(sorted(vals.tolist()))
    # equal predictions keep their input order (stable), unlike Lua's table.sort whose order among ties is unspecified
    idx = np.arange(7, dtype=np.float32); tagged = images.copy(); tagged[:, 1, 0, 0] = idx
    best, _ = nn_utils.sortImagesByPrediction(d, tagged, False, 2, 4)
    assert [int(b[1, 0, 0]) for b in best] == [1, 3]


def test_mode_switches():
    g, d = FakeG(), FakeD()
    nn_utils.switchToEvaluationMode(g, d); assert (g.mode, d.mode) == ("eval", "eval")
    nn_utils.switchToTrainingMode(g, d); assert (g.mode, d.mode) == ("train", "train")


def test_checkpoint_round_trip_and_mismatch(tmp_path):
    g, d = FakeG(), FakeD()
    g.p += 0.5; g.run *= 3; d.p -= 2
    path = str(tmp_path / "adversarial.npz")
    checkpoint.save(path, g, d, epoch=12, opt={"batchSize": 128, "colorSpace": "rgb"})
    g2, d2 = FakeG(), FakeD()
    epoch, opt = checkpoint.load(path, g2, d2)
    assert epoch == 12 and int(opt["batchSize"]) == 128 and str(opt["colorSpace"]) == "rgb"
    assert np.array_equal(g2.p, g.p) and np.array_equal(g2.run, g.run) and np.array_equal(d2.p, d.p)
    other = FakeG(); other.C = 1
    with pytest.raises(ValueError):
        checkpoint.load(path, other, FakeD())
    short = FakeD(); short.nparams = 4
    before = g2.p.copy()
    with pytest.raises(ValueError):
        checkpoint.load(path, g2, short)
    assert np.array_equal(g2.p, before)              # nothing was loaded on a mismatch


def test_rate_with_V_list_and_tensor():
    """NN_UTILS.rateWithV (utils/nn_utils.lua:686-711): 1 - mean of the first SoftMax column, for a list of images or one tensor."""
    class FakeV:
        def forward(self, x):
            p = x[:, 0, 0, 0].astype(np.float32)                       # P(fake) = the image's first pixel
            return np.stack([p, 1 - p], 1)
    rng = np.random.default_rng(1)
    x = rng.uniform(0, 1, (9, 3, 32, 32)).astype(np.float32)
    want = 1.0 - float(x[:, 0, 0, 0].astype(np.float64).mean())
    assert abs(nn_utils.rateWithV(FakeV(), x) - want) < 1e-6
    assert abs(nn_utils.rateWithV(FakeV(), [x[i] for i in range(9)]) - want) < 1e-6


def test_torch7_checkpoint_round_trip_with_real_layouts(tmp_path):
    """saveAs / torch.load through the Torch7 binary format (train.lua:252-261,127-137) with stand-in networks that have the real
    parameter counts: the module trees must take and give back getParameters()'s vector exactly, and refuse another architecture."""
    from catgen import lib as cl
    from oracle import pyoracle as po

    class Net:
        def __init__(self, kind, C, n, nrun, seed):
            r = np.random.default_rng(seed)
            self.kind, self.C, self.nz, self.nparams = kind, C, 100, n
            self.p, self.run = r.standard_normal(n).astype(np.float32), r.uniform(0.5, 1.5, nrun).astype(np.float32)
        def get_params(self): return self.p.copy()
        def set_params(self, a): self.p = np.asarray(a, np.float32).copy()
        def get_bn_running(self): return self.run.copy()
        def set_bn_running(self, a): self.run = np.asarray(a, np.float32).copy()

    g, d = Net(cl.G32UPC, 3, 5191687, 2 * (512 + 256 + 128), 1), Net(cl.D32_ST3, 3, 6664777, 0, 2)
    path = str(tmp_path / "adversarial.net")
    checkpoint.save_torch7(path, g, d, epoch=4, opt={"batchSize": 128, "colorSpace": "rgb"}, normalize_mean=0.1)
    g2, d2 = Net(cl.G32UPC, 3, 5191687, 2 * (512 + 256 + 128), 3), Net(cl.D32_ST3, 3, 6664777, 0, 4)
    rest = checkpoint.load_torch7(path, g2, d2)
    assert rest["epoch"] == 4 and rest["opt"]["colorSpace"] == "rgb" and rest["normalize_mean"] == 0.1
    assert np.array_equal(g2.p, g.p) and np.array_equal(d2.p, d.p) and np.array_equal(g2.run, g.run)
    other = Net(cl.G32UP, 3, 2470406, 2 * (256 + 128), 5)
    before = other.p.copy()
    with pytest.raises(ValueError):
        checkpoint.load_torch7(path, other, None)
    assert np.array_equal(other.p, before)                              # nothing was loaded
    assert po.V32_nparams(3) == 6288258


def test_bench_layer_names_cover_the_step():
    """bench.py labels the per-shape roofline rows by algorithmic MFLOP per launch: layers that do the same amount of work share a row and its label names both."""
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
    names = b.conv_shape_names(128)
    assert len(names) == 22 and len(set(names.values())) == 22
    assert names[round(2.0 * 128 * 32 * 32 * 128 * 256 * 25 / 1e6)].startswith("G.conv3")
    ambiguous = [v for v in names.values() if " or " in v]              # two pairs of layers do the same work: the label says so instead of guessing
    assert len(ambiguous) == 2 and any("G.conv1" in v and "G.conv2" in v for v in ambiguous)
