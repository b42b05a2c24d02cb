"""`persistent_class` with the attributes callers rely on (torch_utils/persistence.py:35):
`init_args` / `init_kwargs` so that `type(G)(*G.init_args, **G.init_kwargs)` re-instantiates a generator
(viz/renderer.py:199).  Source embedding for pickles (persistence.py:99-122) is out of scope."""

import copy


def persistent_class(orig_class):
    class Decorator(orig_class):
        def __init__(self, *args, **kwargs):
            self._init_args = copy.deepcopy(args)
            self._init_kwargs = copy.deepcopy(kwargs)
            super().__init__(*args, **kwargs)

        @property
        def init_args(self):
            return copy.deepcopy(self._init_args)

        @property
        def init_kwargs(self):
            return copy.deepcopy(self._init_kwargs)

    Decorator.__name__ = orig_class.__name__
    Decorator.__qualname__ = orig_class.__qualname__
    Decorator.__module__ = orig_class.__module__
    return Decorator
