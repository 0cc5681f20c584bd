"""patch() completeness, checked without a GPU: every helper a HotPath method calls on ``self`` must exist on a patched
class that had none of them (round-1 bug: ``_extend_state`` was left behind and the second planning step raised)."""
import ast
import inspect
import textwrap

import algp_b200
from algp_b200 import agent as A


def test_patch_installs_every_hotpath_member():
    class Blank(object):
        pass
    A.patch(Blank)
    for name, member in A.HotPath.__dict__.items():
        if name.startswith("__"):
            continue
        assert Blank.__dict__[name] is member, name
    assert isinstance(Blank.__dict__["cov_matrix"], property)
    assert algp_b200.patch is A.patch


def test_every_self_method_call_in_hotpath_resolves_after_patch():
    class Blank(object):
        pass
    A.patch(Blank)
    tree = ast.parse(textwrap.dedent(inspect.getsource(A.HotPath)))
    called = set()
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and \
                isinstance(node.func.value, ast.Name) and node.func.value.id == "self":
            called.add(node.func.attr)
    assert "_extend_state" in called and "_state_for" in called
    missing = sorted(name for name in called if not hasattr(Blank, name))
    assert not missing, "HotPath calls self.%s() but patch() does not install it" % missing
