"""Minimal stand-in for gin-config, only so the *reference* (/root/reference) can be imported
in this container to pin the oracle.  Test infrastructure; not part of the product.

`bind('NerfMLP', net_depth=8)` registers kwargs that `@gin.configurable` classes receive at
construction, which is all the reference's hot path needs from gin.
"""
_BINDINGS = {}


def clear_config():
    _BINDINGS.clear()


def bind(name, **kw):
    _BINDINGS.setdefault(name, {}).update(kw)


def bindings(name):
    return dict(_BINDINGS.get(name, {}))


def configurable(obj=None, **_unused):
    def wrap(o):
        if isinstance(o, type):
            orig, name = o.__init__, o.__name__

            def init(self, *a, __orig=orig, __name=name, **k):
                __orig(self, *a, **{**_BINDINGS.get(__name, {}), **k})

            o.__init__ = init
        return o

    if obj is None or not callable(obj):
        return wrap
    return wrap(obj)


def add_config_file_search_path(_p):
    pass


def parse_config_files_and_bindings(*_a, **_k):
    pass


def operative_config_str():
    return ''


def config_str():
    return ''


def config_scope(_name):
    import contextlib
    return contextlib.nullcontext()
