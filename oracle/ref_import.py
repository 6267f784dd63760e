"""Import the UNMODIFIED reference (minfenli/refnerf-pl) from /root/reference for oracle pinning.

TEST INFRASTRUCTURE ONLY.  Nothing under `refnerf_pl_b200/` may import this module; it is used by
`oracle/make_golden.py` (fixture generation) and by `tests/` that pin `oracle/refnerf_oracle.py`
against the reference when /root/reference is present (it is absent on the GPU box).

Recipe: stub packages for gin / dm_pix / lpips / pycolmap on sys.path, the
`numpy.math = math` shim (ref_utils.py:55 uses np.math.factorial), and the flat `Name.param = value`
lines of the shipped gin files parsed into constructor kwargs.
"""
import ast
import math
import os
import re
import sys

REFERENCE_ROOT = os.environ.get('REFNERF_REFERENCE_ROOT', '/root/reference')
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'stubs')


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'internal'))


def parse_gin_file(path):
    """Flat literal bindings `Scope.param = value` -> {scope: {param: value}}."""
    out = {}
    with open(path) as f:
        text = f.read().replace('\\\n', ' ')
    for line in text.splitlines():
        line = line.split('#', 1)[0].strip()
        m = re.match(r'^([A-Za-z_][\w]*)\.([A-Za-z_][\w]*)\s*=\s*(.+)$', line)
        if not m:
            continue
        scope, key, val = m.groups()
        try:
            val = ast.literal_eval(val.strip())
        except Exception:
            continue
        out.setdefault(scope, {})[key] = val
    return out


def load(gin_file='blender_refnerf.gin', extra_bindings=None):
    """Returns (modules namespace, Config instance) for the reference with the gin file applied."""
    if not available():
        raise RuntimeError('reference tree not present at ' + REFERENCE_ROOT)
    import numpy
    if not hasattr(numpy, 'math'):
        numpy.math = math
    for p in (REFERENCE_ROOT, _STUBS):
        if p not in sys.path:
            sys.path.insert(0, p)
    import gin  # the stub
    gin.clear_config()
    b = parse_gin_file(os.path.join(REFERENCE_ROOT, 'configs', gin_file)) if gin_file else {}
    for scope, kw in (extra_bindings or {}).items():
        b.setdefault(scope, {}).update(kw)
    for scope, kw in b.items():
        gin.bind(scope, **kw)
    import warnings
    warnings.filterwarnings('ignore')
    from internal import configs, coord, image, math as rmath, models, ref_utils, render, stepfun, utils
    from internal import train_utils

    class NS:
        pass

    ns = NS()
    ns.configs, ns.coord, ns.image, ns.math, ns.models = configs, coord, image, rmath, models
    ns.ref_utils, ns.render, ns.stepfun, ns.utils, ns.train_utils = ref_utils, render, stepfun, utils, train_utils
    ns.bindings = b
    config = configs.Config(**{k: v for k, v in b.get('Config', {}).items()
                               if k in configs.Config.__dataclass_fields__})
    return ns, config
