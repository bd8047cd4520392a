"""Minimal name -> class registry with the reference's interface (lbasicsr/utils/registry.py:11-47),
used only when the reference package is not importable."""


class Registry:
    def __init__(self, name):
        self._name = name
        self._obj_map = {}

    def _do_register(self, name, obj):
        if name in self._obj_map:
            raise AssertionError(f"An object named '{name}' was already registered in '{self._name}' registry!")
        self._obj_map[name] = obj

    def register(self, obj=None):
        if obj is None:
            def deco(func_or_class):
                self._do_register(func_or_class.__name__, func_or_class)
                return func_or_class
            return deco
        self._do_register(obj.__name__, obj)
        return obj

    def get(self, name):
        if name not in self._obj_map:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
        return self._obj_map[name]

    def __contains__(self, name):
        return name in self._obj_map

    def keys(self):
        return self._obj_map.keys()


ARCH_REGISTRY = Registry("arch")


def build_network(opt):
    """Same call shape as lbasicsr/archs/__init__.py:19-28: opt = {'type': 'SAVSR', **kwargs}."""
    opt = dict(opt)
    return ARCH_REGISTRY.get(opt.pop("type"))(**opt)
