"""Name -> class registries, as Detectron2's (fvcore ``Registry``): the yaml names a component
(``MODEL.META_ARCHITECTURE``, ``MODEL.PROPOSAL_GENERATOR.NAME``, ``MODEL.ROI_HEADS.NAME``, ``MODEL.BACKBONE.NAME``), importing
the module that defines it registers it (reference train_net.py:14-20, "hacky way to register")."""


class Registry:
    def __init__(self, name):
        self._name = name
        self._map = {}

    def register(self, obj=None, name=None):
        def deco(o):
            key = name or o.__name__
            if key in self._map and self._map[key] is not o:
                raise KeyError("An object named '{}' was already registered in '{}' registry!".format(key, self._name))
            self._map[key] = o
            return o
        return deco if obj is None else deco(obj)

    def get(self, name):
        if name not in self._map:
            raise KeyError("No object named '{}' found in '{}' registry!".format(name, self._name))
        return self._map[name]

    def __contains__(self, name):
        return name in self._map

    def __iter__(self):
        return iter(self._map.items())


META_ARCH_REGISTRY = Registry("META_ARCH")
BACKBONE_REGISTRY = Registry("BACKBONE")
PROPOSAL_GENERATOR_REGISTRY = Registry("PROPOSAL_GENERATOR")
ROI_HEADS_REGISTRY = Registry("ROI_HEADS")
