"""A small continuous experiment file in the reference's format (cf. its
experiments/ddpg_move_to_a_point_time_unit.py): the env grid is what the
sweep front end reads."""
from ray import tune
from collections import OrderedDict

num_seeds = 2

var_env_configs = OrderedDict(
    {
        "state_space_dim": [2],
        "action_space_dim": [2],
        "delay": [0],
        "make_denser": [True],
        "transition_noise": [0],
        "reward_noise": [0],
        "target_point": [[0, 0]],
        "target_radius": [0.5],
        "state_space_max": [10],
        "action_space_max": [1],
        "action_loss_weight": [0.0],
        "time_unit": [0.2, 1.0, 4.0],
        "transition_dynamics_order": [1, 2],
        "dummy_seed": [i for i in range(num_seeds)],
    }
)

var_configs = OrderedDict({"env": var_env_configs})

env_config = {
    "env": "RLToy-v0",
    "horizon": 100,
    "env_config": {
        "seed": 0,
        "state_space_type": "continuous",
        "action_space_type": "continuous",
        "inertia": 1,
        "reward_scale": 1.0,
        "reward_function": "move_to_a_point",
    },
}

algorithm = "DDPG"
