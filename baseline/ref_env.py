"""The UNMODIFIED reference (ShioMisaka/fce-yolo = ultralytics 8.3.242 fork) as an installed package under
``baseline/_ref`` - measurement and test infrastructure, never imported by ``fce_yolo_b200/``.

``baseline/_ref`` is git-ignored but travels to the GPU box with the snapshot (like the built ``.so``).  It is made by

    pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>

(``--no-deps``: the image already holds torch / torchvision / numpy / opencv / PyYAML ...; only dependency RESOLUTION
fails offline.  The copy under /tmp is needed because the build writes ``*.egg-info`` next to the sources and
``/root/reference`` is read-only.)  ``__graft_entry__.build()`` calls :func:`ensure_installed` in the build container.

Three users:
  * ``bench.py --impl reference``: ``YOLO(cfg).predict(tensor, device="cpu")`` as shipped;
  * ``bench.py``'s ``gpu_baseline`` leg (``baseline/ref_gpu_baseline.py``): the reference's own DetectionModel through
    stock PyTorch (cuDNN / cuBLAS, bf16 channels_last, eager and ``torch.compile``) on the same GPU - SURVEY 2.3's
    "kernel to beat";
  * ``tests/test_gpu_reference_dropin.py``: ``fce_yolo_b200.install()`` on the reference's own classes, driven through
    the reference's own ``YOLO(...).predict`` / ``.val`` on the GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_DIR = os.path.join(HERE, "_ref")
REF_SRC = "/root/reference"


def installed() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "ultralytics", "__init__.py"))


def ensure_installed(verbose: bool = False) -> bool:
    """Installs the reference into baseline/_ref when its source tree is present (build container).  Returns whether
    baseline/_ref is usable afterwards.  On the GPU box the prebuilt directory is used as it arrived."""
    if installed():
        return True
    if not os.path.isdir(os.path.join(REF_SRC, "ultralytics")):
        return False
    tmp = tempfile.mkdtemp(prefix="fce_ref_src_")
    try:
        src = os.path.join(tmp, "src")
        shutil.copytree(REF_SRC, src, symlinks=True, ignore=shutil.ignore_patterns(".git", "docs", "examples"))
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
               "/opt/wheelhouse", "--target", REF_DIR, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            print(r.stdout[-2000:], r.stderr[-2000:])
        return r.returncode == 0 and installed()
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def import_reference():
    """Makes ``import ultralytics`` resolve to baseline/_ref and returns the module.  Raises RuntimeError when the
    reference is not installed (callers decide whether that is a skip or an ``unavailable`` line)."""
    if not installed():
        raise RuntimeError("reference not installed under baseline/_ref (run __graft_entry__.build() where "
                           "/root/reference exists)")
    os.environ.setdefault("YOLO_CONFIG_DIR", os.path.join(tempfile.gettempdir(), "fce_ulcfg"))
    os.makedirs(os.environ["YOLO_CONFIG_DIR"], exist_ok=True)
    os.environ.setdefault("YOLO_VERBOSE", "false")
    os.environ.setdefault("YOLO_OFFLINE", "true")
    sys.dont_write_bytecode = True
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import ultralytics

    if not os.path.abspath(ultralytics.__file__).startswith(REF_DIR):
        raise RuntimeError(f"another ultralytics is shadowing baseline/_ref: {ultralytics.__file__}")
    return ultralytics


def reference_model(yaml_name: str, variant=None, seed: int | None = 0):
    """The reference's own DetectionModel (tasks.py:339) built by its own parse_model from its own YAML, with the
    variant rows of the BASELINE configs substituted as a cfg dict (tasks.py:377), fused (tasks.py:223-252) and -
    unless seed is None - filled with the seeded synthetic weights every parity test uses."""
    import_reference()
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from ultralytics.nn.tasks import DetectionModel, yaml_model_load

    from fce_yolo_b200.tasks import variant_cfg
    from fce_yolo_b200.weights import load_synthetic

    d = variant_cfg(yaml_model_load(yaml_name), variant)
    m = DetectionModel(d, verbose=False).eval()
    m.fuse(verbose=False)
    if seed is not None:
        load_synthetic(m, seed)
    return m


def reference_yolo(yaml_name: str, variant=None, seed: int | None = 0):
    """``YOLO(yaml)`` (engine/model.py:81, models/yolo/model.py:26) whose ``.model`` is :func:`reference_model`."""
    import_reference()
    from ultralytics import YOLO
    from ultralytics.cfg import DEFAULT_CFG_DICT

    y = YOLO(yaml_name, task="detect", verbose=False)
    m = reference_model(yaml_name, variant, seed)
    m.args = {**DEFAULT_CFG_DICT, **y.overrides}  # what Model._new attaches (engine/model.py:253-255)
    m.task = "detect"
    y.model = m
    return y


def offline_val_env():
    """What the reference's ``YOLO.val()`` needs to run offline in this image: it dies in ``check_det_dataset`` ->
    ``check_font`` (data/utils.py:475, utils/checks.py:308-336) because matplotlib is not installed and no font can be
    downloaded.  A zero-byte ``Arial.ttf`` in the config directory is found first (checks.py:322-325), and a stub
    ``matplotlib.font_manager`` satisfies the import.  TEST / MEASUREMENT infrastructure only."""
    import types

    import_reference()
    d = os.path.join(os.environ["YOLO_CONFIG_DIR"], "Ultralytics")
    os.makedirs(d, exist_ok=True)
    open(os.path.join(d, "Arial.ttf"), "a").close()
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        m, fm = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.font_manager")
        fm.findSystemFonts = lambda *a, **k: []
        m.font_manager = fm
        sys.modules["matplotlib"], sys.modules["matplotlib.font_manager"] = m, fm
