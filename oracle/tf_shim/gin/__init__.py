"""gin-config stand-in: the decorators are identity (no dependency injection is
configured anywhere on the path exercised here)."""


def _decorator(fn_or_name=None, *args, **kwargs):
  if callable(fn_or_name) and not args and not kwargs:
    return fn_or_name
  return lambda fn: fn


configurable = _decorator
register = _decorator
external_configurable = lambda fn, *a, **k: fn
REQUIRED = object()


def parse_config(*args, **kwargs):
  raise NotImplementedError('gin shim: no configuration support')


def unlock_config():
  import contextlib
  return contextlib.nullcontext()
