"""Import the reference's own Python (`conerf.*` under /root/reference) in this container, where most
of its third-party dependencies (omegaconf, plyfile, pytorch3d, trimesh, tensorboardX, visdom, lpips,
imageio, open3d, matplotlib, cv2 ...) and its CUDA extensions are not importable.

Every module that cannot be found is replaced by a permissive stub, so the reference's module-level
imports succeed and its pure-torch methods (GaussianSplatModel.*, SlaveGaussianSplatTrainer.*,
MasterGaussianSplatTrainer.*) can be EXECUTED unmodified on CPU tensors.  Used only by the golden-vector
generators under tools/ (the vectors are committed under tests/golden/; nothing at test or run time
imports /root/reference).
"""
import importlib
import importlib.abc
import importlib.machinery
import importlib.util
import sys
import types

REFERENCE_ROOT = "/root/reference"
# third-party packages / CUDA extensions of the reference that are absent here, and the reference's own
# root-level `eval.py` / `utils.py` (entry-point plumbing that drags in the evaluators)
STUBBED_TOP_LEVEL = {
    "omegaconf", "plyfile", "pytorch3d", "trimesh", "tensorboardX", "visdom", "lpips", "imageio", "open3d",
    "matplotlib", "cv2", "kornia", "simple_knn", "diff_gaussian_rasterization", "fused_ssim", "eval", "utils",
    "PIL", "sklearn_extra", "easydict", "roma", "pycolmap", "hloc", "mpl_toolkits", "skimage", "torchvision",
    "pytorch_msssim", "nerfacc", "tinycudann", "wandb", "yaml_include",
}


class _AnyMeta(type):
    def __getattr__(cls, name):          # class-level access: OmegaConf.register_new_resolver(...)
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()


class _Anything(metaclass=_AnyMeta):
    """Callable, subclassable, attribute-transparent placeholder."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __iter__(self):
        return iter(())


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return type(name, (_Anything,), {})


class _StubLoader(importlib.abc.Loader):
    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []          # behaves as a package: sub-modules resolve to stubs too
        return m

    def exec_module(self, module):
        pass


class _StubFinder(importlib.abc.MetaPathFinder):
    """Appended LAST to sys.meta_path: only reached when no real finder has the module."""

    def __init__(self):
        self.stubbed = []

    def find_spec(self, fullname, path, target=None):
        top = fullname.split(".")[0]
        if top not in STUBBED_TOP_LEVEL:
            return None          # anything else that is missing is a genuine ImportError
        self.stubbed.append(fullname)
        return importlib.machinery.ModuleSpec(fullname, _StubLoader(), is_package=True)


_finder = None


def enable():
    """Makes `import conerf...` work; returns the list that collects the names of stubbed modules."""
    global _finder
    if _finder is None:
        _finder = _StubFinder()
        sys.meta_path.append(_finder)
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
    return _finder.stubbed


def reference_classes():
    """(GaussianSplatModel, SlaveGaussianSplatTrainer, MasterGaussianSplatTrainer, master module)."""
    enable()
    gsm = importlib.import_module("conerf.model.gaussian_fields.gaussian_splat_model")
    slave = importlib.import_module("conerf.trainers.slave_gaussian_trainer")
    master = importlib.import_module("conerf.trainers.master_gaussian_trainer")
    return gsm.GaussianSplatModel, slave.SlaveGaussianSplatTrainer, master.MasterGaussianSplatTrainer, master
