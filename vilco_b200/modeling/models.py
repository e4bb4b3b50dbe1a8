"""Registries / factories with the reference's names (MQ/libs/modeling/models.py:1-60)."""
backbones, necks, generators, meta_archs = {}, {}, {}, {}


def _reg(table):
    def register(name):
        def decorator(cls):
            table[name] = cls
            return cls
        return decorator
    return register


register_backbone, register_neck = _reg(backbones), _reg(necks)
register_generator, register_meta_arch = _reg(generators), _reg(meta_archs)


def make_backbone(name, **kwargs):
    return backbones[name](**kwargs)


def make_neck(name, **kwargs):
    return necks[name](**kwargs)


def make_meta_arch(name, **kwargs):
    return meta_archs[name](**kwargs)


def make_generator(name, **kwargs):
    return generators[name](**kwargs)
