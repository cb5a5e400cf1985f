"""Host-side behaviour of the fluxion mirror: Chain editing, context plumbing, adapters, LoRA.

Modelled on the reference's own unit tests (tests/fluxion/layers/test_chain.py,
tests/adapters/test_adapter.py, test_lora.py, test_range_adapter.py) - same behaviours, fresh
tests.  CPU only."""

import pytest
import torch

import refiners_b200.fluxion.layers as fl
from refiners_b200.fluxion.adapters import Adapter, Conv2dLora, LinearLora, Lora, LoraAdapter, auto_attach_loras
from refiners_b200.fluxion.context import ContextProvider
from refiners_b200.fluxion.utils import no_grad


def test_unique_child_names_are_state_dict_keys():
    chain = fl.Chain(fl.Linear(2, 2), fl.SiLU(), fl.Linear(2, 2), fl.Chain(fl.Linear(2, 3)))
    assert list(chain._modules) == ["Linear_1", "SiLU", "Linear_2", "Chain"]
    assert set(chain.state_dict()) == {
        "Linear_1.weight", "Linear_1.bias", "Linear_2.weight", "Linear_2.bias", "Chain.Linear.weight", "Chain.Linear.bias",
    }
    assert chain.Linear_2 is chain[2] and chain["Chain"] is chain[3]


def test_chain_threads_tuples_and_kinds():
    with no_grad():
        x = torch.randn(2, 4)
        par = fl.Parallel(fl.Identity(), fl.Multiply(2.0))
        a, b = par(x)
        assert torch.equal(a, x) and torch.equal(b, 2 * x)
        dist = fl.Distribute(fl.Multiply(3.0), fl.Multiply(-1.0))
        c, d = fl.Chain(par, dist)(x)
        assert torch.equal(c, 3 * x) and torch.equal(d, -2 * x)
        assert torch.equal(fl.Sum(fl.Identity(), fl.Multiply(2.0))(x), 3 * x)
        assert torch.equal(fl.Residual(fl.Multiply(2.0))(x), 3 * x)
        assert fl.Concatenate(fl.Identity(), fl.Identity(), dim=1)(x).shape == (2, 8)
        assert torch.equal(fl.Passthrough(fl.Multiply(5.0))(x)[0], x)
        with pytest.raises(Exception):
            dist(x)  # arity mismatch


def test_find_walk_layers():
    inner = fl.Chain(fl.Linear(1, 1), fl.Chain(fl.Linear(1, 1)))
    top = fl.Chain(fl.SiLU(), inner, fl.Linear(1, 1))
    assert len(list(top.layers(fl.Linear))) == 3
    assert len(list(top.layers(fl.Chain))) == 1            # does not descend below a match
    assert len(list(top.layers(fl.Chain, recurse=True))) == 2
    assert top.find(fl.SiLU) is top[0]
    assert top.find(fl.GroupNorm) is None
    assert top.find_parent(inner[0]) is inner
    assert top.ensure_find_parent(top[2]) is top
    with pytest.raises(AssertionError):
        top.ensure_find(fl.GroupNorm)
    with pytest.raises(ValueError):
        list(top.walk(list[int]))  # subscripted generics are rejected
    assert top.layer(("Chain", "Chain", "Linear"), fl.Linear) is inner[1][0]

    def prune(m, parent):
        if m is inner:
            raise StopIteration
        return isinstance(m, fl.Linear)

    assert [m for m, _ in top.walk(prune)] == [top[2]]


def test_insert_replace_remove_pop_keep_parents_and_keys():
    a, b, c = fl.Chain(), fl.Chain(), fl.Chain()
    top = fl.Chain(a, b)
    assert a.parent is top and b.parent is top
    top.insert(1, c)
    assert list(top) == [a, c, b] and c.parent is top
    assert list(top._modules) == ["Chain_1", "Chain_2", "Chain_3"]
    top.remove(c)
    assert c.parent is None and list(top._modules) == ["Chain_1", "Chain_2"]
    popped = top.pop(0)
    assert popped is a and a.parent is None and list(top._modules) == ["Chain"]
    d = fl.Chain()
    top.replace(b, d)
    assert b.parent is None and d.parent is top
    top.append(fl.SiLU())
    top.insert_before_type(fl.SiLU, fl.ReLU())
    top.insert_after_type(fl.SiLU, fl.Sigmoid())
    assert [type(m).__name__ for m in top] == ["Chain", "ReLU", "SiLU", "Sigmoid"]
    with pytest.raises(ValueError):
        top.insert_before_type(fl.GroupNorm, fl.ReLU())
    with pytest.raises(ValueError):
        top.remove(fl.ReLU())
    with pytest.raises(IndexError):
        top.pop(10)
    with pytest.raises(ValueError):
        top.foo = fl.ReLU()  # modules cannot be attached by attribute


def test_context_is_shared_and_reset_after_forward():
    class Producer(fl.Chain):
        def init_context(self):
            return {"store": {"items": []}}

    sink = fl.SetContext("store", "items", callback=lambda items, x: items.append(x.sum().item()))
    source = fl.UseContext("store", "items").compose(lambda items: torch.tensor(items))
    top = Producer(fl.Chain(sink), fl.Chain(fl.Chain(source)))
    with no_grad():
        out = top(torch.ones(3))
    assert out.tolist() == [3.0]
    # init_context is re-applied after every forward (reference chain.py:256)
    assert top.provider.get_context("store")["items"] == []
    top.set_context("extra", {"k": 1})
    assert source.use_context("extra") == {"k": 1}  # propagated to nested providers
    provider = ContextProvider.create({"a": {"x": 1}})
    provider.update_contexts({"a": {"y": 2}, "b": {"z": 3}})
    assert provider.get_context("a") == {"x": 1, "y": 2} and provider.get_context("b") == {"z": 3}


def test_structural_copy_shares_leaves_not_chains():
    lin = fl.Linear(2, 2)
    src = fl.Chain(fl.Chain(lin, fl.SiLU()), fl.UseContext("c", "k"))
    dup = src.structural_copy()
    assert dup is not src and dup[0] is not src[0]
    assert dup[0][0] is lin                       # weights are shared
    assert dup[0].parent is dup and src[0].parent is src
    assert dup[1].context == "c" and dup[1].key == "k"
    sliced = src[0:1]
    assert len(sliced) == 1 and sliced[0][0] is lin


def test_chain_error_reports_tree_and_inputs():
    bad = fl.Chain(fl.Linear(4, 4), fl.Chain(fl.Linear(3, 3)))
    with no_grad(), pytest.raises(fl.ChainError) as info:
        bad(torch.randn(2, 4))
    text = str(info.value)
    assert ">>>" in text and "Linear" in text and "shape=(2, 4)" in text


def test_repr_tree_format():
    chain = fl.Chain(fl.Linear(1, 1, bias=False), fl.Linear(1, 1, bias=False), fl.Residual(fl.SiLU()))
    text = repr(chain)
    assert text.splitlines()[0] == "(CHAIN)"
    assert "Linear(in_features=1, out_features=1, device=cpu, dtype=float32) (x2)" in text
    assert "(RES) Residual()" in text and "└── SiLU()" in text


class _Wrap(fl.Chain, Adapter[fl.Chain]):
    def __init__(self, target: fl.Chain) -> None:
        with self.setup_adapter(target):
            super().__init__(target)


def test_adapter_inject_eject_roundtrip():
    target = fl.Chain(fl.Linear(2, 2))
    parent = fl.Chain(target)
    before = repr(parent)
    adapter = _Wrap(target)
    assert target.parent is parent                # untouched until inject
    adapter.inject()
    assert parent[0] is adapter and target.parent is adapter and adapter.parent is parent
    adapter.eject()
    assert parent[0] is target and target.parent is parent and adapter.parent is None
    assert repr(parent) == before
    outer = _Wrap(target)
    outer.inject()
    inner = _Wrap(target)
    inner.inject()
    assert parent[0] is outer and outer[0] is inner and inner[0] is target
    outer.eject()
    assert parent[0] is inner
    inner.eject()
    assert repr(parent) == before
    with pytest.raises(AssertionError):
        class _Bad(fl.Module, Adapter[fl.Chain]):  # adapters must be Chains
            pass


def test_lora_algebra_and_manager():
    torch.manual_seed(0)
    lora = LinearLora("a", in_features=8, out_features=6, rank=2, scale=2.0)
    assert torch.count_nonzero(lora.up.weight) == 0                  # up starts at zero
    assert lora.scale == 2.0 and lora.ensure_find(fl.Multiply).scale == 2.0
    lora.scale = 0.5
    assert lora.ensure_find(fl.Multiply).scale == 0.5
    twin = Lora.from_weights("b", down=lora.down.weight, up=torch.randn(6, 2))
    assert isinstance(twin, LinearLora) and twin.rank == 2 and twin.in_features == 8
    conv = Lora.from_weights("c", down=torch.randn(4, 3, 3, 3), up=torch.randn(5, 4, 1, 1))
    assert isinstance(conv, Conv2dLora) and conv.kernel_size == (3, 1) and conv.padding == (1, 0)
    with pytest.raises(ValueError):
        Lora.from_weights("d", down=torch.randn(2, 2), up=torch.randn(2, 2, 1, 1))

    base = fl.Linear(8, 6)
    holder = fl.Chain(base)
    twin.up.weight.data.normal_()
    adapter = LoraAdapter(base, twin).inject(holder)
    x = torch.randn(3, 8)
    with no_grad():
        want = base(x) + twin.scale * (x @ twin.down.weight.T @ twin.up.weight.T)
        assert torch.allclose(holder(x), want, atol=1e-6)
    assert adapter.names == ["b"] and list(holder._modules) == ["LoraAdapter"]
    adapter.add_lora(LinearLora("z", in_features=8, out_features=6, rank=1))
    assert adapter.names == ["b", "z"] and adapter.scales == {"b": 1.0, "z": 1.0}
    adapter.remove_lora("z")
    adapter.eject()
    assert holder[0] is base


def test_auto_attach_respects_include_exclude_and_sanity_check():
    class Block(fl.Chain):
        pass

    model = fl.Chain(Block(fl.Linear(4, 4), fl.Linear(4, 4)), fl.Chain(fl.Linear(4, 4)))
    loras = {f"l{i}": LinearLora("n", in_features=4, out_features=4, rank=2) for i in range(2)}
    failed = auto_attach_loras(loras, model, include=["Block"])
    assert failed == []
    assert len(list(model.layers(LoraAdapter, recurse=True))) == 2
    assert isinstance(model[1][0], fl.Linear)            # outside the included block: untouched
    extra = {"x": LinearLora("n", in_features=4, out_features=4, rank=2)}
    with pytest.raises(ValueError):
        auto_attach_loras(extra, model, include=["Block"])  # same name everywhere already


def test_state_dict_roundtrip_through_safetensors(tmp_path):
    from refiners_b200.fluxion.utils import save_to_safetensors

    a = fl.Chain(fl.Linear(3, 3), fl.LayerNorm(3))
    b = fl.Chain(fl.Linear(3, 3), fl.LayerNorm(3))
    path = tmp_path / "w.safetensors"
    save_to_safetensors(path, a.state_dict())
    b.load_from_safetensors(path)
    assert all(torch.equal(x, y) for x, y in zip(a.state_dict().values(), b.state_dict().values()))
    with pytest.raises(RuntimeError):
        fl.Chain(fl.Linear(3, 3)).load_from_safetensors(path)  # strict by default


def test_slicing_and_shape_ops():
    x = torch.arange(24.0).reshape(2, 3, 4)
    assert torch.equal(fl.Slicing(dim=2, start=1, end=3)(x), x[:, :, 1:3])
    assert torch.equal(fl.Slicing(dim=1, start=-2)(x), x[:, -2:])
    assert fl.Slicing(dim=1, start=5)(x).shape == (2, 0, 4)
    assert fl.Reshape(4, 3)(x).shape == (2, 4, 3)
    assert fl.Flatten(1)(x).shape == (2, 12)
    assert fl.Unflatten(1)(fl.Flatten(1)(x), torch.Size((3, 4))).shape == (2, 3, 4)
    assert fl.Permute(2, 0, 1)(x).shape == (4, 2, 3)
    assert torch.equal(fl.Multiply(2.0, 1.0)(x), 2 * x + 1)
    assert fl.GetArg(1)(x, x + 1)[0, 0, 0] == 1


def test_sdpa_module_matches_reference_formula_and_slicing():
    torch.manual_seed(0)
    q, k, v = torch.randn(2, 10, 16), torch.randn(2, 7, 16), torch.randn(2, 7, 16)
    with no_grad():
        full = fl.ScaledDotProductAttention(num_heads=4)(q, k, v)
        plain = fl.ScaledDotProductAttention(num_heads=4, is_optimized=False)(q, k, v)
        sliced = fl.ScaledDotProductAttention(num_heads=4, slice_size=3)(q, k, v)
    assert torch.allclose(full, plain, atol=1e-6) and torch.allclose(full, sliced, atol=1e-6)


def test_value_epoch_counts_real_changes_only():
    """The CUDA-graph runner re-captures when a public scalar of any module changes (engine/graph.py): re-assigning the value a
    module already holds - a pipeline that sets ``adapter.scale = s`` on every step - must not count."""
    import refiners_b200.fluxion.layers as fl
    from refiners_b200.fluxion.layers.base import value_epoch

    m = fl.Multiply(scale=0.5)
    start = value_epoch()
    m.scale = 0.5
    assert value_epoch() == start
    m.scale = 0.25
    assert value_epoch() == start + 1
    m.scale = 0.25
    m._private = 3
    assert value_epoch() == start + 1
    m.scale = 1  # an int is not the float 1.0 held before? it is a different value here anyway
    assert value_epoch() == start + 2
    m.scale = 1.0  # same number, different type: counts (kernels may specialise on it)
    assert value_epoch() == start + 3

    class Forwarding(fl.Chain):  # an adapter-style property that forwards to a child
        @property
        def scale(self) -> float:
            return self[0].scale

        @scale.setter
        def scale(self, value: float) -> None:
            self[0].scale = value

    f = Forwarding(fl.Multiply(scale=2.0))
    start = value_epoch()
    f.scale = 2.0
    assert value_epoch() == start
    f.scale = 3.0
    assert value_epoch() == start + 2 and f[0].scale == 3.0  # the property assignment and the child's attribute
