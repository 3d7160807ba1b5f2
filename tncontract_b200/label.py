"""String labels with a prime history (host metadata only; no device work).

Behavioural mirror of the reference's label module (label.py:17-92): a primed
label remembers the label it was made from, so un-priming returns the very
same object and ``noprime_label`` recovers the original plain string.
"""
import uuid

__all__ = ['prime_label', 'unprime_label', 'noprime_label', 'prime_level', 'unique_label']


class Label(str):
    """A ``str`` that knows its parent label (label.py:17-54)."""

    def __new__(cls, value, **kwargs):
        return super().__new__(cls, value)

    def __init__(self, label, parent=None):
        if parent is None and isinstance(label, Label):
            parent = label._parent
        self._parent = parent

    @property
    def parent(self):
        return self._parent

    @property
    def origin(self):
        # walk up until something that has no ``parent`` attribute (a plain
        # str -- or None for a Label that was never derived from another)
        node = self
        while hasattr(node, "parent"):
            node = node.parent
        return node

    @property
    def parents(self):
        depth, node = 0, self
        while hasattr(node, "parent") and node.parent is not None:
            node = node.parent
            depth += 1
        return depth


def prime_label(label, prime="'"):
    """label.py:57-59."""
    return Label(str(label) + prime, parent=label)


def unprime_label(label, prime="'"):
    """label.py:62-71: returns the parent object; ValueError when not primed."""
    if not hasattr(label, "parent"):
        raise ValueError("label is not primed")
    parent = label.parent
    if str(parent) + prime != label:
        raise ValueError("label is not primed with \"" + prime + "\"")
    return parent


def noprime_label(label):
    """label.py:74-79."""
    return label.origin if hasattr(label, "origin") else label


def prime_level(label):
    """label.py:82-87."""
    return label.parents if hasattr(label, "parents") else 0


def unique_label():
    """label.py:90-92."""
    return str(uuid.uuid4())
