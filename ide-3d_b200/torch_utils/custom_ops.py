"""Plugin loader with the reference's interface (torch_utils/custom_ops.py:59 `get_plugin`).

The reference JIT-compiles a pybind11 module per op with torch.utils.cpp_extension.load.  Here every
"plugin" is a thin Python object over the prebuilt C-ABI library (ide-3d_b200/lib/libide3d_b200.so, built
by ide-3d_b200/build.py), so get_plugin() resolves names instead of compiling sources.  It keeps the
reference's contract: cached by module name, prints according to `verbosity`, raises on failure.
"""

from .. import _lib, _plugins

verbosity = 'brief'  # 'none' | 'brief' | 'full'   (custom_ops.py:24)

_cached_plugins = dict()


def get_plugin(module_name, sources=None, headers=None, source_dir=None, **build_kwargs):
    assert verbosity in ['none', 'brief', 'full']
    if module_name in _cached_plugins:
        return _cached_plugins[module_name]
    if verbosity == 'full':
        print(f'Setting up PyTorch plugin "{module_name}"...')
    elif verbosity == 'brief':
        print(f'Setting up PyTorch plugin "{module_name}"... ', end='', flush=True)
    try:
        _lib.get_lib()                                  # raises if the CUDA library was not built
        module = _plugins.PLUGINS[module_name]
    except Exception:
        if verbosity == 'brief':
            print('Failed!')
        raise
    if verbosity == 'full':
        print(f'Done setting up PyTorch plugin "{module_name}".')
    elif verbosity == 'brief':
        print('Done.')
    _cached_plugins[module_name] = module
    return module
