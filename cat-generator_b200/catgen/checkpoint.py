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


# ---------------------------------------------------------------- Torch7 .net files (train.lua:252-261 writes, :127-137 and :119-123 read)
def save_torch7(path, MODEL_G, MODEL_D, epoch=0, opt=None, plot_data=None, normalize_mean=None, normalize_std=None):
    """saveAs(filename) (train.lua:252-261): `{D=, G=, opt=, plot_data=, epoch=, normalize_mean=, normalize_std=}` in Torch7's binary
    serialisation, the networks as nn module trees shaped like models.lua's constructors with weights that are views into one flat
    storage each (what getParameters() leaves behind).  See catgen/torch7.py for the format and its parity-unpinned status."""
    from . import torch7, lib
    obj = {"D": torch7.tree_D(MODEL_D.C, MODEL_D.get_params()),
           "G": torch7.tree_G(MODEL_G.kind == lib.G32UPC, MODEL_G.C, MODEL_G.nz, MODEL_G.get_params(), MODEL_G.get_bn_running()),
           "opt": dict(opt or {}), "plot_data": plot_data if plot_data is not None else {}, "epoch": int(epoch)}
    if normalize_mean is not None:
        obj["normalize_mean"] = normalize_mean
    if normalize_std is not None:
        obj["normalize_std"] = normalize_std
    torch7.save(path, obj)


def load_torch7(path, MODEL_G=None, MODEL_D=None, MODEL_V=None):
    """torch.load of a reference checkpoint (train.lua:127-137: tmp.G / tmp.D / tmp.epoch; :119-123: tmp.V).  Each network present in
    both the file and the arguments receives getParameters()'s vector (and the BatchNormalization running statistics); a length
    mismatch raises ValueError before anything is loaded.  Returns the table without the networks (epoch, opt, plot_data, ...)."""
    from . import torch7
    tab = torch7.load(path)
    if not isinstance(tab, dict):
        raise ValueError("%s does not hold a table" % path)
    todo = []
    for key, model in (("G", MODEL_G), ("D", MODEL_D), ("V", MODEL_V)):
        if model is None or key not in tab:
            continue
        flat, run = torch7.flat_parameters(tab[key]), torch7.bn_running(tab[key])
        if flat.size != model.nparams:
            raise ValueError("%s in %s has %d parameters, the network here has %d" % (key, path, flat.size, model.nparams))
        if run.size != model.get_bn_running().size:
            raise ValueError("%s in %s has %d running statistics, the network here has %d" % (key, path, run.size, model.get_bn_running().size))
        todo.append((model, flat, run))
    for model, flat, run in todo:
        model.set_params(flat)
        if run.size:
            model.set_bn_running(run)
    return {k: v for k, v in tab.items() if k not in ("G", "D", "V")}
