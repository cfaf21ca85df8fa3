"""Golden-vector cases shared by make_golden.py (reference side) and the
parity tests (oracle / CUDA side).  Shapes follow BASELINE.json's configs and
the reference's own experiments/tests:

  c1_*        example.py:48-69 / experiments' 8x8 toy (BASELINE config #1)
  c2_*        dqn_seq_del.py / dqn_p_r_noises.py shape (config #2)
  custom_8x5  tests/test_mdp_playground.py:1999-2011 style custom P/R (8x5)
  c3_*        sac_move_to_a_point_p_order_2.py / ddpg_..._irr_dims.py (#3)
  c4_*        dqn_image_representations_sh_quant.py (#4)
"""
import copy

import numpy as np

LANE_SEED = 4242

_D = dict(seed=0, state_space_type="discrete", action_space_type="discrete",
          state_space_size=8, action_space_size=8, reward_density=0.25,
          terminal_state_density=0.25, generate_random_mdp=True,
          completely_connected=True, dummy_seed=3)

_C = dict(seed=0, state_space_type="continuous",
          action_space_type="continuous", state_space_dim=6,
          action_space_dim=6, relevant_indices=[0, 1],
          irrelevant_features=True, transition_dynamics_order=2, inertia=1.0,
          time_unit=0.5, target_radius=0.05, target_point=[0.0, 0.0],
          state_space_max=10.0, action_space_max=1.0,
          reward_function="move_to_a_point")

_rngP = np.random.default_rng(123)
_P85 = _rngP.integers(0, 8, size=(8, 5))
_P85[6, :] = 6
_P85[7, :] = 7
_R85 = np.round(_rngP.normal(size=(8, 5)), 3)

_G = dict(seed=0, state_space_type="grid", grid_shape=(8, 8), delay=0,
          sequence_length=1, reward_function="move_to_a_point",
          target_point=[5, 5], make_denser=True)

CASES = {
    "c1_seq1": dict(config=dict(_D, sequence_length=1, delay=0)),
    "c2_seq3_del2_noise": dict(config=dict(
        _D, sequence_length=3, delay=2, transition_noise=0.1,
        reward_noise=0.25)),
    "c2_every1": dict(config=dict(
        _D, sequence_length=3, delay=2, transition_noise=0.1,
        reward_noise=0.25, reward_every_n_steps=True)),
    "seq2_denser": dict(config=dict(
        _D, seed=6, sequence_length=2, make_denser=True,
        reward_every_n_steps=1, delay=1)),
    "diam3_seq4": dict(config=dict(
        _D, seed=3, diameter=3, sequence_length=4, delay=1,
        reward_every_n_steps=1, transition_noise=0.05)),
    "rdist_scale_shift": dict(config=dict(
        _D, seed=7, sequence_length=2, reward_dist=[0.1, 1.0],
        reward_scale=2.5, reward_shift=-1.75, term_state_reward=-3.0,
        reward_noise=0, reward_every_n_steps=1)),
    "repeats_seq3": dict(config=dict(
        _D, seed=5, sequence_length=3, repeats_in_sequences=True,
        reward_every_n_steps=1)),
    "notmax_diam2": dict(config=dict(
        _D, seed=11, maximally_connected=False, sequence_length=2,
        diameter=2, delay=8, reward_every_n_steps=1)),
    "big50": dict(config=dict(
        _D, seed=4, action_space_size=50, state_space_size=50,
        sequence_length=2, delay=4, transition_noise=0.25, reward_noise=1.0,
        reward_every_n_steps=1), steps=60, horizon=20),
    "custom_8x5": dict(config=dict(
        seed=0, state_space_type="discrete", action_space_type="discrete",
        state_space_size=8, action_space_size=5, use_custom_mdp=True,
        transition_function=_P85, reward_function=_R85,
        init_state_dist=np.array([1 / 6] * 6 + [0, 0]),
        terminal_states=[6, 7], delay=1, reward_noise=0.1,
        transition_noise=0.2, reward_scale=0.5)),
    # discrete irrelevant_features (dqn_irr_dims.py shape: sizes [8, 8]; plus
    # unequal sub-spaces, diameter, noise, images of both sub-states)
    "irr_8x8_noise": dict(config=dict(
        _D, seed=8, state_space_size=[8, 8], action_space_size=[8, 8],
        irrelevant_features=True, sequence_length=2, delay=1,
        transition_noise=0.1, reward_noise=0.5, reward_every_n_steps=1)),
    "irr_6x10_diam2": dict(config=dict(
        _D, seed=5, state_space_size=[12, 20], action_space_size=[6, 10],
        irrelevant_features=True, diameter=2, sequence_length=3,
        transition_noise=0.3, reward_every_n_steps=1)),
    "irr_8x5_det": dict(config=dict(
        _D, seed=3, state_space_size=[8, 5], action_space_size=[8, 5],
        irrelevant_features=True)),
    "irr_img_all": dict(config=dict(
        _D, seed=2, state_space_size=[8, 8], action_space_size=[8, 8],
        irrelevant_features=True, transition_noise=0.1,
        image_representations=True,
        image_transforms="shift,scale,rotate,flip", image_sh_quant=2,
        image_ro_quant=1, image_scale_range=(0.5, 1.5)),
        lanes=2, steps=24, horizon=8),
    # grid envs (tests/test_mdp_playground.py:792-1190 shapes)
    "grid_dense_term": dict(config=dict(
        _G, reward_scale=3.0, term_state_reward=-0.25,
        terminal_states=[[5, 5], [2, 3], [2, 4], [3, 3], [3, 4]])),
    "grid_sparse_noise": dict(config=dict(
        _G, seed=4, make_denser=False, transition_noise=0.3, reward_noise=1.0,
        reward_scale=2.0, reward_shift=0.5, term_state_reward=2.0)),
    "grid_irr_5x9": dict(config=dict(
        _G, seed=5, grid_shape=(5, 9), target_point=[1, 7],
        irrelevant_features=True, transition_noise=0.2,
        reward_every_n_steps=2)),
    "grid_img": dict(config=dict(
        _G, seed=1, image_representations=True, reward_scale=2.0,
        terminal_states=[[5, 5], [2, 3], [2, 4]]),
        lanes=2, steps=24, horizon=8),
    "grid_img_irr": dict(config=dict(
        _G, seed=2, grid_shape=(4, 6), target_point=[1, 2],
        irrelevant_features=True, image_representations=True,
        image_width=60, image_height=48, transition_noise=0.25),
        lanes=2, steps=24, horizon=8),
    "c4_img_shift": dict(config=dict(
        _D, seed=2, sequence_length=1, image_representations=True,
        image_transforms="shift", image_sh_quant=4, image_width=100,
        image_height=100), lanes=3, steps=30, horizon=10),
    "c4_img_all": dict(config=dict(
        _D, seed=2, sequence_length=1, image_representations=True,
        image_transforms="shift,scale,rotate,flip", image_sh_quant=2,
        image_ro_quant=1, image_scale_range=(0.5, 1.5)),
        lanes=3, steps=40, horizon=10),
    "img_none_64x48": dict(config=dict(
        _D, seed=2, sequence_length=1, image_representations=True,
        image_width=64, image_height=48), lanes=2, steps=12, horizon=6),
    "c3_order2": dict(config=dict(_C)),
    "cont_order1_clip": dict(config=dict(
        _C, transition_dynamics_order=1, time_unit=1.0, state_space_max=2.0,
        target_radius=0.5)),
    "cont_order3": dict(config=dict(
        _C, transition_dynamics_order=3, time_unit=0.3, inertia=2.0,
        state_space_max=3.0)),
    "cont_noise_delay": dict(config=dict(
        _C, transition_noise=0.05, reward_noise=0.1, delay=2,
        reward_scale=2.0, reward_shift=0.5, action_loss_weight=0.3,
        state_space_max=1.5)),
    "cont_sparse": dict(config=dict(
        _C, make_denser=False, target_radius=1.0, state_space_max=2.0,
        term_state_reward=5.0)),
    "cont_term_boxes": dict(config=dict(
        _C, state_space_dim=2, action_space_dim=2, irrelevant_features=False,
        state_space_max=5.0, terminal_states=[[1.0, 1.0], [-2.0, 0.5]],
        term_state_edge=1.5, time_unit=1.0)),
    "cont_unbounded": dict(config={
        k: v for k, v in _C.items()
        if k not in ("state_space_max", "action_space_max")}),
    # per-dimension inertia (rl_toy_env.py:519-537, :1654): a list / float64
    # array makes `action / inertia` float64 (the top derivative leaves fp32),
    # an array of dtype_s keeps the division in fp32
    "cont_inertia_list": dict(config=dict(
        _C, transition_dynamics_order=3, time_unit=0.4,
        inertia=[1.0, 2.0, 0.5, 3.0, 1.5, 0.7], state_space_max=4.0)),
    "cont_inertia_f32": dict(config=dict(
        _C, inertia=np.array([1.0, 2.0, 0.5, 3.0, 1.5, 0.7], dtype=np.float32),
        state_space_max=4.0)),
    # fp32 env, default (float64) target_point, dense reward through the delay
    # FIFO: the reference keeps Python floats in reward_buffer
    "cont_delay_default_target": dict(config={
        k: v for k, v in dict(_C, state_space_dim=2, action_space_dim=2,
                              irrelevant_features=False, delay=3,
                              reward_noise=0.2, reward_scale=1.5,
                              state_space_max=2.0).items()
        if k not in ("target_point", "relevant_indices")}),
    # move_along_a_line (rl_toy_env.py:1865-1910, :2546-2576): reward = - mean
    # distance of the last sequence_length relevant states from their fitted line
    "cont_line_seq10": dict(config=dict(
        seed=0, state_space_type="continuous", action_space_type="continuous",
        state_space_dim=4, action_space_dim=4, transition_dynamics_order=1,
        inertia=1.0, time_unit=1.0, delay=0, sequence_length=10,
        reward_scale=1.0, reward_function="move_along_a_line",
        action_space_max=1.0), steps=60, horizon=30),
    "cont_line_seq3_delay": dict(config=dict(
        seed=3, state_space_type="continuous", action_space_type="continuous",
        state_space_dim=6, action_space_dim=6, relevant_indices=[0, 2, 3],
        irrelevant_features=True, transition_dynamics_order=2, inertia=2.0,
        time_unit=0.5, delay=2, sequence_length=3, reward_scale=2.0,
        reward_shift=-0.5, reward_noise=0.1, transition_noise=0.02,
        reward_function="move_along_a_line", state_space_max=3.0,
        action_space_max=1.0, terminal_states=[[2.5, 2.5, 2.5]],
        term_state_edge=1.0), steps=60, horizon=20),
    "cont_line_2d": dict(config=dict(
        seed=5, state_space_type="continuous", action_space_type="continuous",
        state_space_dim=2, action_space_dim=2, transition_dynamics_order=1,
        time_unit=0.25, sequence_length=4, reward_every_n_steps=2,
        reward_function="move_along_a_line", state_space_max=2.0,
        action_space_max=1.0), steps=60, horizon=30),
    # round 2: longer / wider replays of the two headline shapes
    "c2_big": dict(config=dict(
        _D, sequence_length=3, delay=2, transition_noise=0.1,
        reward_noise=0.25, reward_every_n_steps=True),
        lanes=32, steps=200, horizon=25),
    "c3_big": dict(config=dict(_C, transition_noise=0.02, reward_noise=0.05),
                   lanes=32, steps=200, horizon=40),
    "cont_img": dict(config=dict(
        _C, state_space_dim=4, action_space_dim=4,
        image_representations=True, state_space_max=5.0,
        terminal_states=[[1.0, 1.0], [-2.0, 0.5]], term_state_edge=1.5,
        time_unit=1.0, target_point=[3.0, -3.0]),
        lanes=2, steps=24, horizon=8),
}


def materialise(config):
    return copy.deepcopy(config)
