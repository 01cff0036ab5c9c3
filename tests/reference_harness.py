"""Runs the reference's OWN scripts, unmodified, inside the test process.

The scripts import packages this image does not have (matplotlib, seaborn, pyproj, osgeo, geopandas, shapely; SURVEY.md
§0.3) and plot / block on plt.show().  `stubs()` installs inert stand-ins for those modules for the duration of a `with`
block — every attribute is a callable that returns another inert object — except pyproj, whose Transformer is a real
EPSG:4326 <-> EPSG:32650 transform (ransac_b200.geo, Krueger series), because read_camera_locations needs it.

The reference sources are read from where they lie: /root/reference in the build container, or the unmodified staging
copy `baseline/_ref/` (git-ignored, written by __graft_entry__.build(); the GPU box has no /root/reference), or
$B2R_REFERENCE_DIR.  Nothing is copied into the tracked tree."""
import contextlib
import importlib.util
import os
import runpy
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_dir():
    for d in (os.environ.get("B2R_REFERENCE_DIR"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if d and os.path.isfile(os.path.join(d, "main_v1.py")) and os.path.isfile(os.path.join(d, "testpro-K.py")):
            return d
    return None


class _Inert:
    """Absorbs any use: attribute access, calls, indexing, iteration (as a pair: `fig, ax = plt.subplots()`), context."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Inert()

    def __call__(self, *a, **k):
        return _Inert()

    def __getitem__(self, k):
        return _Inert()

    def __setitem__(self, k, v):
        pass

    def __iter__(self):
        return iter((_Inert(), _Inert()))

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def __len__(self):
        return 0

    def __bool__(self):
        return False


def _inert_module(name):
    m = types.ModuleType(name)
    m.__getattr__ = lambda attr: _Inert()      # PEP 562: any attribute of the module
    m.__path__ = []                            # so that `import a.b` treats it as a package
    return m


class _Transformer:
    """pyproj.Transformer.from_crs between EPSG:4326 and EPSG:32650 (main_v1.py:38-39 with always_xy=True: lon, lat;
    the plot helpers at main_v1.py:74, :92 use the authority order of EPSG:4326: lat, lon).  Scalars or arrays."""

    def __init__(self, forward, always_xy):
        self.forward, self.always_xy = forward, always_xy

    @classmethod
    def from_crs(cls, src, dst, always_xy=False):
        s, d = str(src).lower(), str(dst).lower()
        assert {s, d} == {"epsg:4326", "epsg:32650"}, (src, dst)
        return cls(s == "epsg:4326", always_xy)

    def transform(self, x, y):
        import numpy as np
        from ransac_b200 import geo
        scalar = np.ndim(x) == 0
        if self.forward:
            lon, lat = (x, y) if self.always_xy else (y, x)
            a, b = geo.wgs84_to_utm50n(lon, lat)
        else:
            lon, lat = geo.utm50n_to_wgs84(x, y)
            a, b = (lon, lat) if self.always_xy else (lat, lon)
        return (float(a), float(b)) if scalar else (np.asarray(a), np.asarray(b))


STUBBED = ["matplotlib", "matplotlib.pyplot", "matplotlib.font_manager", "mpl_toolkits", "mpl_toolkits.mplot3d", "seaborn",
           "osgeo", "osgeo.gdal", "geopandas", "shapely", "shapely.geometry", "pyproj", "plotly", "plotly.express", "plotly.graph_objects",
           "rasterio"]


@contextlib.contextmanager
def stubs(cv2_module=None):
    """Install the stand-ins (and, optionally, `cv2_module` as sys.modules['cv2']); restore sys.modules afterwards."""
    saved = {k: sys.modules.get(k) for k in STUBBED + (["cv2"] if cv2_module is not None else [])}
    try:
        for name in STUBBED:
            if name == "pyproj":
                m = types.ModuleType("pyproj")
                m.Transformer = _Transformer
            else:
                m = _inert_module(name)
            sys.modules[name] = m
        if cv2_module is not None:
            sys.modules["cv2"] = cv2_module
        yield
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def import_main_v1(workdir, cv2_module=None):
    """`import main_v1` from the reference directory, unmodified.  Import-time side effects: logging.basicConfig writes
    ./debug.log (main_v1.py:33) — hence `workdir` — and two prints.  Returns the module; its `cv2` global is `cv2_module`
    when given (the injection point SURVEY.md §8b names), else the real OpenCV."""
    ref = reference_dir()
    cwd = os.getcwd()
    os.chdir(workdir)
    try:
        with stubs(cv2_module):
            spec = importlib.util.spec_from_file_location("main_v1_reference", os.path.join(ref, "main_v1.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
    finally:
        os.chdir(cwd)
    if cv2_module is not None:
        mod.cv2 = cv2_module
    return mod


def run_testpro_k(workdir, cv2_module=None):
    """Executes testpro-K.py as a script (it runs its pipeline at import, testpro-K.py:236): returns its globals."""
    ref = reference_dir()
    cwd = os.getcwd()
    os.chdir(workdir)
    try:
        with stubs(cv2_module):
            return runpy.run_path(os.path.join(ref, "testpro-K.py"), run_name="testpro_k_reference")
    finally:
        os.chdir(cwd)


VARIANT_SCRIPTS = ["process.py", "testpro.py", "test_pro.py", "test02.py"]


def load_definitions(script, workdir, cv2_module=None):
    """The other pipeline variants of the reference (process.py, testpro.py, test_pro.py, test02.py) run a whole job at
    import time on images this checkout does not have, so they cannot be imported as they are.  This executes, from the
    file where it lies, only the top-level statements that DEFINE things — imports, function and class definitions,
    plain assignments (`grid_code_min = 7`, `geo_transformer = GeoCoordTransformer()`) — and skips the job (`do_it(...)`,
    the `if img == ...` chain, prints, logging.basicConfig).  The function bodies — find_homographies, find_homography,
    estimate_camera_pose: the hot-path callers SURVEY.md §8(a) cites — are the reference's, unmodified."""
    import ast
    ref = reference_dir()
    path = os.path.join(ref, script)
    with open(path, encoding="utf-8") as f:
        tree = ast.parse(f.read(), filename=path)
    keep = (ast.Import, ast.ImportFrom, ast.FunctionDef, ast.ClassDef, ast.Assign, ast.AnnAssign)
    tree.body = [node for node in tree.body if isinstance(node, keep)]
    mod = types.ModuleType("reference_" + script.replace(".py", "").replace("-", "_"))
    mod.__file__ = path
    cwd = os.getcwd()
    os.chdir(workdir)
    try:
        with stubs(cv2_module):
            exec(compile(tree, path, "exec"), mod.__dict__)
    finally:
        os.chdir(cwd)
    if cv2_module is not None:
        mod.cv2 = cv2_module
    return mod
