"""Table-driven constructor arguments for the drop-in modules.

The reference's modules take long positional/keyword argument lists (e.g. `sr.SoftRenderer` has 31).  Here
each module declares ONE table of (name, default) pairs in the reference's positional order and binds
`*args, **kwargs` against it, so the accepted spellings are exactly the reference's while the modules
themselves stay small.
"""


def bind(owner, fields, args, kwargs):
    """Returns {name: value} for `fields` = ((name, default), ...) given positional `args` / keyword `kwargs`."""
    names = [n for n, _ in fields]
    if len(args) > len(names):
        raise TypeError("%s() takes at most %d positional arguments (%d given)" % (owner, len(names), len(args)))
    out = dict(fields)
    for n, v in zip(names, args):
        out[n] = v
    for k, v in kwargs.items():
        if k not in out:
            raise TypeError("%s() got an unexpected keyword argument %r" % (owner, k))
        if k in names[:len(args)]:
            raise TypeError("%s() got multiple values for argument %r" % (owner, k))
        out[k] = v
    return out


def pick(cfg, mapping):
    """Sub-dictionary {new_name: cfg[old_name]} for mapping = ((old_name, new_name), ...)."""
    return {new: cfg[old] for old, new in mapping}
