"""Generates tests/golden/ref_registry_golden.json: a transcript of the REFERENCE's own plugin machinery
(mmdet/utils/registry.py Registry / build_from_cfg, mmdet/models/builder.py build) on a scripted scenario - values,
reprs, exception types and messages - run in the build container where /root/reference exists (mmcv is a
one-function stand-in: is_str).  tests/test_host.py replays the same script on hvrnet_b200.registry / builder.

    python tests/golden/make_registry_golden.py
"""
import importlib.util
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference/mmdet'


def scenario(Registry, build_from_cfg, build):
    """Returns the transcript; used with the reference's objects here and with ours in the test."""
    import torch.nn as nn
    log = []

    def attempt(label, fn):
        try:
            log.append([label, 'ok', repr(fn())])
        except Exception as e:            # noqa: BLE001 - the transcript records type and message
            log.append([label, type(e).__name__, str(e.args[0]) if e.args else ''])
    R = Registry('thing')

    class A(nn.Module):
        def __init__(self, x=0, y=2):
            super().__init__()
            self.x, self.y = x, y

        def __repr__(self):
            return 'A(x=%r, y=%r)' % (self.x, self.y)

    class B(A):
        pass
    attempt('register A', lambda: R.register_module(A).__name__)
    attempt('register B', lambda: R.register_module(B).__name__)
    attempt('repr', lambda: R)
    attempt('name', lambda: R.name)
    attempt('get A', lambda: R.get('A').__name__)
    attempt('get missing', lambda: R.get('C'))
    attempt('module_dict keys', lambda: sorted(R.module_dict.keys()))
    attempt('duplicate', lambda: R.register_module(A))
    attempt('not a class', lambda: R.register_module(3))
    cfg = dict(type='A', x=1)
    attempt('build defaults', lambda: build_from_cfg(cfg, R, dict(x=9, y=5)))
    attempt('cfg untouched', lambda: cfg)
    attempt('build class type', lambda: build_from_cfg(dict(type=B, x=3), R))
    attempt('build unknown', lambda: build_from_cfg(dict(type='C'), R))
    attempt('build bad type', lambda: build_from_cfg(dict(type=3.5), R))
    attempt('build no type', lambda: build_from_cfg(dict(x=1), R))
    attempt('build non-dict', lambda: build_from_cfg([1], R))
    attempt('build bad default_args', lambda: build_from_cfg(dict(type='A'), R, default_args=[1]))
    attempt('build unexpected kwarg', lambda: build_from_cfg(dict(type='A', z=1), R))
    attempt('build list', lambda: type(build([dict(type='A', x=1), dict(type='B')], R)).__name__)
    attempt('build list items', lambda: [m for m in build([dict(type='A', x=1), dict(type='B')], R, dict(y=7))])
    attempt('build single', lambda: build(dict(type='B', x=4), R, dict(y=8)))
    return log


def main():
    mmcv = types.ModuleType('mmcv')
    mmcv.is_str = lambda s: isinstance(s, str)
    sys.modules['mmcv'] = mmcv
    spec = importlib.util.spec_from_file_location('ref_registry', os.path.join(REF, 'utils', 'registry.py'))
    reg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(reg)
    # models/builder.py:8-15 imports the registries of the package; its `build` is cut out and run on the real build_from_cfg
    import ast
    import torch.nn as nn
    src = os.path.join(REF, 'models', 'builder.py')
    node = next(n for n in ast.parse(open(src).read()).body if isinstance(n, ast.FunctionDef) and n.name == 'build')
    ns = dict(nn=nn, build_from_cfg=reg.build_from_cfg)
    exec(compile(ast.Module(body=[node], type_ignores=[]), src, 'exec'), ns)
    log = scenario(reg.Registry, reg.build_from_cfg, ns['build'])
    json.dump(log, open(os.path.join(HERE, 'ref_registry_golden.json'), 'w'), indent=1)
    for row in log:
        print(row)


if __name__ == '__main__':
    main()
