"""Constructor-argument capture for pickling, same contract as learning_to_adapt/utils/serializable.py:16-49."""
import inspect


class Serializable(object):
    def __init__(self, *args, **kwargs):
        self.__args = args
        self.__kwargs = kwargs

    def quick_init(self, locals_):
        if getattr(self, "_serializable_initialized", False):
            return
        spec = inspect.getfullargspec(self.__init__)
        in_order = [locals_[arg] for arg in spec.args][1:]
        varargs = locals_[spec.varargs] if spec.varargs else tuple()
        kwargs = dict(locals_[spec.varkw]) if spec.varkw else dict()
        self.__args = tuple(in_order) + tuple(varargs)
        self.__kwargs = kwargs
        setattr(self, "_serializable_initialized", True)

    def __getstate__(self):
        return {"__args": self.__args, "__kwargs": self.__kwargs}

    def __setstate__(self, d):
        out = type(self)(*d["__args"], **d["__kwargs"])
        self.__dict__.update(out.__dict__)
