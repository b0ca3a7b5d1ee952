"""Import shim for the UNMODIFIED reference at /root/reference (authoring container only).

Used by tests/golden/make_golden.py to produce the committed golden vectors and by the optional
`-m "not gpu"` cross-checks that skip when /root/reference is absent (it does not exist on the GPU
box).  Nothing here is copied from the reference: the three one-line defects documented in
SURVEY.md section 0.2 are repaired by text substitution on the source *as loaded at run time*.

    defect 1  vidExample.py:164  proc_dt[i] is a length-1 array      -> proc_dt[i, 0]
    defect 2  vidExample.py:134  im0 never assigned                  -> im0 = im at end of loop body
    defect 3  utils/KLT.py:87    ([*xy0, dx, dy]).astype(...)        -> (xy0 + [dx, dy]).astype(...)
"""
import inspect
import os
import sys
import types

REF_ROOT = os.environ.get("VELOCITY_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "utils", "KLT.py"))


_loaded = None


def load():
    """Returns a namespace with the reference modules (KLT, NLS, MSV, transforms, common, images)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    if "plots" not in sys.modules:  # bokeh is not installed; presentation only
        stub = types.ModuleType("plots")
        stub.plotresults = lambda *a, **k: None
        stub.imshow = lambda *a, **k: None
        sys.modules["plots"] = stub
    import utils.common as common
    import utils.images as images
    import utils.KLT as KLT
    import utils.MSV as MSV
    import utils.NLS as NLS
    import utils.transforms as transforms

    # defect 3
    src = inspect.getsource(KLT.KLTregional)
    bad = "([*xy0, dx, dy]).astype(np.float32)"
    if bad in src:
        src = src.replace(bad, "(xy0 + [dx, dy]).astype(np.float32)")
        exec(compile(src, "<reference KLTregional, defect 3 repaired>", "exec"), KLT.__dict__)

    ns = types.SimpleNamespace(common=common, images=images, KLT=KLT, MSV=MSV, NLS=NLS, transforms=transforms,
                               root=REF_ROOT)
    _loaded = ns
    return ns


def load_vid_example():
    """Returns the reference's vidExamplefcn with defects 1-2 repaired (run it with cwd=REF_ROOT)."""
    ns = load()
    path = os.path.join(REF_ROOT, "vidExample.py")
    with open(path) as f:
        src = f.read()
    src = src.replace("S[i, :] = (i, proc_dt[i],", "S[i, :] = (i, proc_dt[i, 0],")
    marker = "        im_gaussian = cv2.GaussianBlur(im, (3, 3), 0)"
    assert marker in src
    src = src.replace(marker, "        im0 = im\n" + marker)
    src = src.replace("            del im0\n", "")
    mod = types.ModuleType("vidExample_shimmed")
    mod.__file__ = path
    exec(compile(src, "<reference vidExample, defects 1-2 repaired>", "exec"), mod.__dict__)
    return mod, ns
