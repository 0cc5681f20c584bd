"""algp_b200.patch() installs the accelerated hot path on the reference's OWN Agent class with the
reference's signatures.  Needs /root/reference (build container only); no GPU compute."""
import inspect
import os
import sys

import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "agent.py")), reason="reference not mounted")


def test_patch_keeps_reference_signatures(golden_dir):
    sys.path.insert(0, golden_dir)
    import gpytorch_standin
    gpytorch_standin.install()
    sys.path.insert(0, REF)
    try:
        import agent as ref_agent
    finally:
        sys.path.remove(REF)
    import algp_b200
    before = {n: inspect.signature(getattr(ref_agent.Agent, n)) for n in
              ("greedy", "best_path", "predict", "_post_update", "get_sampled_dataset", "update_model")}
    algp_b200.patch(ref_agent.Agent)
    for n, sig in before.items():
        fn = getattr(ref_agent.Agent, n)
        assert fn is getattr(algp_b200.HotPath, n)
        assert list(inspect.signature(fn).parameters) == list(sig.parameters), n
    assert isinstance(ref_agent.Agent.__dict__["cov_matrix"], property)
    # the drop-in GPR mirrors the reference constructor and public methods
    import models as ref_models
    for n in ("__init__", "fit", "reset", "set_train_data", "cov_mat", "predict", "get_embeddings"):
        assert list(inspect.signature(getattr(algp_b200.GPR, n)).parameters) == \
            list(inspect.signature(getattr(ref_models.GPR, n)).parameters), n
    import utils as ref_utils
    for n in ("entropy_from_cov", "predictive_distribution", "to_torch", "to_numpy"):
        assert list(inspect.signature(getattr(algp_b200.utils, n)).parameters) == \
            list(inspect.signature(getattr(ref_utils, n)).parameters), n
    assert algp_b200.CONST == ref_utils.CONST


def test_patch_env_keeps_reference_signature():
    """algp_b200.paths.patch_env(FieldEnv) replaces get_all_paths with the same parameter list (env.py:197)."""
    import types
    for name in ["seaborn", "ipdb", "matplotlib", "matplotlib.pyplot"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    import networkx
    if not hasattr(networkx, "nx"):
        networkx.nx = networkx
    sys.path.insert(0, REF)
    try:
        import env as ref_env
    finally:
        sys.path.remove(REF)
    from algp_b200 import paths as P
    before = list(inspect.signature(ref_env.FieldEnv.get_all_paths).parameters)
    original = ref_env.FieldEnv.get_all_paths
    try:
        P.patch_env(ref_env.FieldEnv)
        assert list(inspect.signature(ref_env.FieldEnv.get_all_paths).parameters) == before
        assert ref_env.FieldEnv.get_all_paths is not original
    finally:
        ref_env.FieldEnv.get_all_paths = original
