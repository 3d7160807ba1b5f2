"""Persistence / interchange with the reference.

The reference has no file format of its own: its fixture ``tncontract/tests/random_10site_mps.dat`` is a
plain pickle of ``tncontract.onedim.onedim_core.MatrixProductState`` whose site tensors are
``tncontract.tensor.Tensor`` objects with the attributes ``data`` (ndarray) and ``_labels`` (list); the
network keeps ``data`` (object ndarray), ``left_label``, ``right_label``, ``phys_label``.  The classes here use
the same attribute names, and ``Tensor.__getstate__`` / ``__setstate__`` move ``data`` between host ndarray
and device array, so that

* ``load(f)`` reads a pickle written by the reference (class paths ``tncontract.*`` are mapped onto
  ``tncontract_b200.*``; Python-2 pickles via ``encoding="latin1"``) or by this package;
* ``dump(obj, f)`` writes a pickle the REFERENCE can read back (class paths rewritten to ``tncontract.*``,
  array data on the host), ``dump(obj, f, reference_paths=False)`` one for this package only.
"""
import io
import pickle
import pickletools

_PREFIX_REF, _PREFIX_OURS = "tncontract", "tncontract_b200"


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == _PREFIX_REF or module.startswith(_PREFIX_REF + "."):
            module = _PREFIX_OURS + module[len(_PREFIX_REF):]
        return super().find_class(module, name)


def load(file, encoding="latin1"):
    """Unpickle a Tensor / network written by the reference or by dump().  ``file``: binary file object."""
    return _Unpickler(file, encoding=encoding).load()


def loads(data, encoding="latin1"):
    return load(io.BytesIO(data), encoding=encoding)


def dumps(obj, reference_paths=True, protocol=2):
    """Pickle bytes.  Protocol <= 3 names classes in text GLOBAL opcodes ("c<module>\\n<name>\\n"), which is what
    lets the module path be rewritten without touching any length field."""
    if protocol > 3:
        raise ValueError("protocols above 3 use length-prefixed class paths; use protocol 2 or 3")
    raw = pickle.dumps(obj, protocol=protocol)
    if not reference_paths:
        return raw
    out, pos = bytearray(), 0
    needle = ("c" + _PREFIX_OURS).encode()
    for op, arg, at in pickletools.genops(raw):
        if op.name == "GLOBAL" and raw[at:at + len(needle)] == needle:
            out += raw[pos:at] + b"c" + _PREFIX_REF.encode()
            pos = at + len(needle)
    out += raw[pos:]
    return bytes(out)


def dump(obj, file, reference_paths=True, protocol=2):
    file.write(dumps(obj, reference_paths=reference_paths, protocol=protocol))
