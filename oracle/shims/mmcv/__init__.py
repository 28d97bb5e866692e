"""mmcv 1.7.0 stand-in: the registry plumbing the reference's model builder needs."""


class Registry:
    def __init__(self, name):
        self.name = name
        self._module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def _register(obj):
            self._module_dict[name or obj.__name__] = obj
            return obj
        if module is not None:
            return _register(module)
        return _register

    def get(self, key):
        return self._module_dict.get(key)

    def build(self, cfg, default_args=None):
        return build_from_cfg(cfg, self, default_args)


def build_from_cfg(cfg, registry, default_args=None):
    args = dict(cfg)
    if default_args:
        for k, v in default_args.items():
            args.setdefault(k, v)
    obj_type = args.pop("type")
    obj = registry.get(obj_type) if isinstance(obj_type, str) else obj_type
    return obj(**args)


class Config(dict):
    @staticmethod
    def fromfile(path):
        raise NotImplementedError("mmcv.Config shim: not needed on the oracle path")
