"""Minimal stand-in for the `immutabledict` wheel (absent offline).

Test infrastructure only: lets /root/reference's hot-path modules import in
the build container so goldens can be generated. Never on the product path.
"""


class immutabledict(dict):
  """Hashable, write-protected dict."""

  def _ro(self, *a, **k):
    raise TypeError("immutabledict is read-only")

  __setitem__ = __delitem__ = clear = pop = popitem = setdefault = update = _ro

  def __hash__(self):
    return hash(frozenset(self.items()))

  def __reduce__(self):
    return (immutabledict, (dict(self),))
