"""Host-side mirror of the reference's network lookup (the drop-in boundary).

The reference has no registry object: `basicsr/models/archs/__init__.py:9-18` imports every `*_arch.py` file of its
folder and `define_network(opt)` (:43-46) pops `opt['type']` and instantiates the first scanned module attribute of that
name with the remaining keys as kwargs (`dynamic_instantiation`, :21-40).  This module restates exactly that behaviour
over `refid_b200/archs/` so configs written for the reference (`options/**.yml`, key `network_g`) build the B200
backend unchanged, and so the tests can exercise the lookup path without the reference tree.
"""
import importlib
import os

_ARCH_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "archs")


def scan_arch_modules(folder=_ARCH_DIR, package="refid_b200.archs"):
    """Import every `*_arch.py` under `folder` (sorted, so the order is not filesystem-dependent)."""
    names = sorted(os.path.splitext(f)[0] for f in os.listdir(folder) if f.endswith("_arch.py"))
    return [importlib.import_module(f"{package}.{n}") for n in names]


def dynamic_instantiation(modules, cls_type, opt):
    """First module exposing `cls_type` wins; unknown types raise ValueError (reference :34-40)."""
    cls_ = None
    for m in modules:
        cls_ = getattr(m, cls_type, None)
        if cls_ is not None:
            break
    if cls_ is None:
        raise ValueError(f"{cls_type} is not found.")
    return cls_(**opt)


def define_network(opt):
    """`opt` is the `network_g` mapping of an option file; `type` is popped (the caller passes a deepcopy,
    basicsr/models/twoImage_event_recurrent_model.py:24)."""
    network_type = opt.pop("type")
    return dynamic_instantiation(scan_arch_modules(), network_type, opt)


def load_options(path):
    """YAML option file -> dict (the subset of basicsr/utils/options.py:31-95 the hot path consumes)."""
    import yaml
    with open(path) as f:
        return yaml.safe_load(f)
