"""`ConfigMixin` / `register_to_config` stand-ins (see package docstring)."""
import functools
import inspect


class _AttrDict(dict):
    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError:  # deepcopy/pickle probe dunder attributes
            raise AttributeError(key)


class ConfigMixin:
    config = _AttrDict()


def register_to_config(init):
    sig = inspect.signature(init)

    @functools.wraps(init)
    def wrapper(self, *args, **kwargs):
        bound = sig.bind(self, *args, **kwargs)
        bound.apply_defaults()
        cfg = _AttrDict({k: v for k, v in bound.arguments.items() if k != "self"})
        object.__setattr__(self, "config", cfg)
        init(self, *args, **kwargs)

    return wrapper
