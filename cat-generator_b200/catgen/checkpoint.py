"""Flat-parameter checkpoints (SURVEY.md section 8(f), row F2 -- the "at least a flat-parameter export" half).

The reference saves `{D=MODEL_D, G=MODEL_G, opt=OPT, plot_data, epoch}` with torch.save every OPT.saveFreq epochs
(/root/reference/train.lua:252-261) and reloads it with torch.load (:127-137); the optimiser state is NOT part of it.  Torch7's
serialisation format belongs to an un-vendored package and is not reproduced here.  What a checkpoint needs in order to continue
training or to sample is each network's flat parameter vector in nn getParameters() order (the order libcatgen uses) plus G's
batch-norm running statistics; this module writes exactly that to one .npz and restores it.  A maintainer with Torch7 can
produce / consume the same arrays with `MODEL:getParameters()` and the BN modules' running_mean / running_var.
"""
import numpy as np

FORMAT = "catgen-flat-1"


def save(path, MODEL_G, MODEL_D, epoch=0, opt=None):
    """train.lua:252-261.  `opt`: a dict of scalars/strings (the reference stores its whole OPT table)."""
    meta = {"format": FORMAT, "epoch": int(epoch), "G_kind": int(MODEL_G.kind), "C": int(MODEL_G.C), "nz": int(MODEL_G.nz)}
    arrays = {"G_params": MODEL_G.get_params(), "G_bn_running": MODEL_G.get_bn_running(), "D_params": MODEL_D.get_params()}
    for k, v in (opt or {}).items():
        arrays["opt_" + str(k)] = np.asarray(v)
    np.savez(path, **arrays, **{"meta_" + k: np.asarray(v) for k, v in meta.items()})


def load(path, MODEL_G, MODEL_D):
    """train.lua:127-137.  The networks must already exist with the same architecture (models.create_G / create_D); returns
    (epoch, opt dict).  Raises ValueError on any mismatch instead of loading a partial state."""
    with np.load(path, allow_pickle=False) as z:
        if str(z["meta_format"]) != FORMAT:
            raise ValueError("not a %s checkpoint: %s" % (FORMAT, path))
        if int(z["meta_G_kind"]) != int(MODEL_G.kind) or int(z["meta_C"]) != int(MODEL_G.C) or int(z["meta_nz"]) != int(MODEL_G.nz):
            raise ValueError("checkpoint is for another generator (kind %d, C %d, nz %d)" % (int(z["meta_G_kind"]), int(z["meta_C"]), int(z["meta_nz"])))
        gp, gr, dp = z["G_params"], z["G_bn_running"], z["D_params"]
        if gp.size != MODEL_G.nparams or dp.size != MODEL_D.nparams:
            raise ValueError("parameter count mismatch: G %d vs %d, D %d vs %d" % (gp.size, MODEL_G.nparams, dp.size, MODEL_D.nparams))
        MODEL_G.set_params(gp); MODEL_G.set_bn_running(gr); MODEL_D.set_params(dp)
        opt = {k[4:]: z[k][()] if z[k].shape == () else z[k] for k in z.files if k.startswith("opt_")}
        return int(z["meta_epoch"]), opt
