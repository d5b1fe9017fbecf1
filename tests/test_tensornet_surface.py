"""CPU: the TensorNet_Dist wrapper surface (tensornet.py:163-217) -- what it accepts, what it refuses before any GPU call."""
import pytest
import torch

from distmlip_b200.implementations.matgl import CHGNet_Dist, TensorNet_Dist
from distmlip_b200.random_init import RandomTensorNet


def test_from_existing_keeps_the_model_attributes_and_a_state_dict_snapshot():
    m = RandomTensorNet(seed=1)
    d = TensorNet_Dist.from_existing(m)
    assert d.dist_enabled is False and d.cutoff == 5.0 and d.units == 64 and d.element_types == m.element_types
    assert set(d._state_dict) == set(m.state_dict())
    assert all(v.dtype == torch.float32 for v in d._state_dict.values())
    with pytest.raises(ValueError):
        TensorNet_Dist.from_existing(m, dtype=torch.float64)
    with pytest.raises(NotImplementedError):
        d.predict_structure_dist(None)
    assert issubclass(TensorNet_Dist, CHGNet_Dist.__mro__[1])  # same engine-backed base as CHGNet_Dist


@pytest.mark.parametrize("change, exc", [
    (dict(is_intensive=True), NotImplementedError),               # tensornet.py:139-142
    (dict(rbf_type="SphericalBessel"), NotImplementedError),
    (dict(activation_type="tanh"), NotImplementedError),
    (dict(equivariance_invariance_group="O(2)"), NotImplementedError),
])
def test_unsupported_configurations_are_refused_before_the_engine_is_created(change, exc):
    m = RandomTensorNet(seed=1)
    for k, v in change.items():
        setattr(m, k, v)
    with pytest.raises(exc):
        TensorNet_Dist.from_existing(m).enable_distributed_mode([0])


def test_cpu_partitions_and_empty_lists_are_refused():
    d = TensorNet_Dist.from_existing(RandomTensorNet(seed=1))
    with pytest.raises(RuntimeError):
        d.enable_distributed_mode(["cpu", "cpu"])
    with pytest.raises(ValueError):
        d.enable_distributed_mode([])
