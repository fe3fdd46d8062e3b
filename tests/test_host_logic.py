"""CPU tests of the host-side mirror of the ReaK::kte modelling API and its chain compiler."""
import numpy as np
import pytest

from reak_b200 import _abi, kte, presets


def test_crs_descriptor_shape():
    s = presets.make("crs6")
    c = kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs)
    assert (c.dim, c.n_coords, c.n_inputs) == (3, 6, 6)
    kinds = [e.kind for e in c.elements]
    # CRS_A465_models.cpp:748-788 chain order: actuator, rotor inertia, joint, link, link inertia
    assert kinds[:5] == [_abi.ACTUATOR_GEN, _abi.INERTIA_GEN, _abi.REVOLUTE_3D, _abi.RIGID_LINK_3D, _abi.INERTIA_3D]
    inertias = [e for e in c.elements if e.kind == _abi.INERTIA_3D]
    assert [e.upstream for e in inertias] == [(1 << (k + 1)) - 1 for k in range(6)]
    acts = [e for e in c.elements if e.kind == _abi.ACTUATOR_GEN]
    assert [a.aux for a in acts] == list(range(6))
    assert all(c.elements[a.frame_b].kind == _abi.REVOLUTE_3D and c.elements[a.frame_b].coord == a.coord for a in acts)
    assert abs(c.desc.base.acceleration[2] - 9.81) < 1e-15 and c.desc.base.position[1] == -3.3


def test_state_layout_follows_dofs_order():
    s = presets.make("crs3")
    s.dofs_gen.reverse()
    s.mass_calc.mCoords.reverse()
    c = kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs)
    joints = [e for e in c.elements if e.kind == _abi.REVOLUTE_3D]
    assert [j.coord for j in joints] == [2, 1, 0]


def test_unsupported_chains_raise():
    s = presets.make("crs3")
    with pytest.raises(kte.UnsupportedChain):
        kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen[:-1], s.inputs)
    with pytest.raises(kte.UnsupportedChain):
        kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs[:-1])
    s2 = presets.make("crs3")
    s2.chain << kte.kte_map("mystery")
    with pytest.raises(kte.UnsupportedChain):
        kte.compile_chain(s2.chain, s2.mass_calc, s2.dofs_gen, s2.inputs)
    s3 = presets.make("crs3")
    s3.chain << kte.rigid_link_2D("flat", kte.frame_2D(), kte.frame_2D(), kte.pose_2D())
    with pytest.raises(kte.UnsupportedChain):
        kte.compile_chain(s3.chain, s3.mass_calc, s3.dofs_gen, s3.inputs)
    with pytest.raises(TypeError):
        kte.mass_matrix_calc() << kte.pose_3D()    # only inertias, generalized coordinates and free-joint frames register


def test_axis_angle_quaternion():
    """core/kinetostatics/unit_test_rotations.cpp:257-330: 45 degrees about z."""
    q = kte.axis_angle_quat(np.pi / 4, (0.0, 0.0, 2.0))
    assert np.allclose(q, [np.cos(np.pi / 8), 0.0, 0.0, np.sin(np.pi / 8)], atol=1e-16)


def test_propagator_argument_checks():
    from reak_b200 import kte_batch_propagator
    p = kte_batch_propagator(presets.make("crs6"))
    assert (p.get_state_dimensions(), p.get_input_dimensions(), p.get_output_dimensions()) == (12, 6, 0)
    assert p.is_serial()
    with pytest.raises(IndexError):
        p.get_state_derivatives(np.zeros((3, 11)), np.zeros((3, 6)))
    with pytest.raises(IndexError):
        p.get_state_derivatives(np.zeros((3, 12)), np.zeros((2, 6)))
    with pytest.raises(IndexError):
        p.get_state_derivatives(np.zeros((3, 12)))  # inputs are required when the chain has actuators
    from reak_b200.propagator import impossible_integration
    with pytest.raises(impossible_integration):
        p.get_next_states(np.zeros((3, 12)), np.zeros((3, 6)), 0.0, 1)
    with pytest.raises(IndexError):
        p.steer_batch(np.zeros((2, 12)), np.zeros((2, 12)), np.zeros((2, 5, 4)))


def test_caller_supplied_result_buffers_are_validated():
    """out= / status= go to the C-ABI by raw pointer: a wrong dtype, shape or layout must be refused on the host
    (the checks run before the library is called, so no device is needed)."""
    from reak_b200 import kte_batch_propagator
    p = kte_batch_propagator(presets.make("crs6"))
    x, u = np.zeros((8, 12)), np.zeros((8, 6))
    for call in (lambda **k: p.get_state_derivatives(x, u, **k), lambda **k: p.get_next_states(x, u, 1e-3, 1, **k),
                 lambda **k: p.get_next_states_multi(x, u, 1e-3, 1, **k),
                 lambda **k: p.rollout(x, np.zeros((8, 2, 6)), 1e-3, 1, **k)):
        with pytest.raises(TypeError):
            call(out=np.zeros((8, 12), dtype=np.float32))
        with pytest.raises(IndexError):
            call(out=np.zeros((7, 12)))
        with pytest.raises(IndexError):
            call(out=np.zeros((8, 13)))
        with pytest.raises(TypeError):
            call(out=np.zeros((12, 8)).T)                      # right shape, not C-contiguous
        with pytest.raises(TypeError):
            call(out=np.zeros((8, 24))[:, ::2])                # strided view
        with pytest.raises(TypeError):
            call(status=np.zeros(8, dtype=np.int64))
        with pytest.raises(IndexError):
            call(status=np.zeros(9, dtype=np.int32))
        with pytest.raises(TypeError):
            call(out=[[0.0] * 12] * 8)                         # not an array at all


def test_inertia_registered_with_mass_calc_but_missing_from_the_chain_is_refused():
    """M comes from mass_calc (kte_nl_system.hpp:271), the forces from the chain: an inertia only mass_calc knows
    would change the reference's M and silently vanish here."""
    s = presets.make("crs3")
    ghost_frame = kte.joint_dependent_frame_3D(kte.frame_3D())
    s.mass_calc << kte.inertia_3D("ghost", ghost_frame, 2.0, (0.1, 0.0, 0.0, 0.1, 0.0, 0.1))
    with pytest.raises(kte.UnsupportedChain, match="not in the chain"):
        kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs)


def test_gen_elements_lowering():
    """rigid_link_gen / spring_gen / damper_gen (rigid_link.cpp:30-75, spring.cpp:32-96, damper.cpp:32-68): anchors that are
    not system dofs become auxiliary coordinates, declared once, right before their first use, with the values they hold"""
    s = presets.make("crs3_gen")
    c = kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs)
    kinds = [e.kind for e in c.elements]
    decl = [e for e in c.elements if e.kind == _abi.COORD_GEN]
    assert [e.coord for e in decl] == [3, 4, 5] and c.n_coords == 3
    assert [list(e.p)[:2] for e in decl] == [[0.25, 0.1], [0.0, 0.0], [-0.4, 0.0]]
    first_use = kinds.index(_abi.SPRING_GEN)
    assert kinds[first_use - 1] == _abi.COORD_GEN and c.elements[first_use].aux == 3
    link = [e for e in c.elements if e.kind == _abi.RIGID_LINK_GEN][0]
    assert (link.coord, link.aux, link.p[0]) == (2, 4, 0.35)
    # a link that ends on a state would overwrite it every doMotion: rejected
    s.chain << kte.rigid_link_gen("bad", s.dofs_gen[0], s.dofs_gen[1], 0.1)
    with pytest.raises(kte.UnsupportedChain):
        kte.compile_chain(s.chain, s.mass_calc, s.dofs_gen, s.inputs)
    # the library validates the same rules on a raw descriptor
    import ctypes as C
    lib = _abi.load_library()
    h = C.c_void_p()
    assert lib.rkb_chain_create(C.byref(c.desc), C.byref(h)) == 0
    assert lib.rkb_chain_is_serial(h) == 0 and lib.rkb_chain_state_dim(h) == 6
    lib.rkb_chain_destroy(h)
    bad = [e for e in c.elements if e.kind == _abi.SPRING_GEN][0]
    keep = bad.aux
    bad.aux = 9  # an auxiliary coordinate nobody declared
    assert lib.rkb_chain_create(C.byref(c.desc), C.byref(h)) == _abi.ERR_INVALID
    bad.aux = keep
