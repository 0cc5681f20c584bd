"""CPU oracle for the algp GP / information-gain hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``algp_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, as the checker or as the
timed CPU baseline, never as the product path.
"""
from .oracle import *  # noqa: F401,F403
