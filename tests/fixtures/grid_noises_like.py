"""A small grid-world experiment file in the reference's format (the
reference ships none for state_space_type "grid"; the keys are those of its
grid example, example.py grid_image_representations_example and
rl_toy_env.py:780-812): the env grid is what the sweep front end reads."""
from ray import tune
from collections import OrderedDict

num_seeds = 2

var_env_configs = OrderedDict(
    {
        "grid_shape": [(8, 8)],
        "target_point": [[5, 5]],
        "make_denser": [True, False],
        "transition_noise": [0, 0.25, 0.5],
        "reward_noise": [0, 1.0],
        "dummy_seed": [i for i in range(num_seeds)],
    }
)

var_configs = OrderedDict({"env": var_env_configs})

env_config = {
    "env": "RLToy-v0",
    "horizon": 50,
    "env_config": {
        "seed": 0,
        "state_space_type": "grid",
        "reward_function": "move_to_a_point",
        "delay": 0,
        "sequence_length": 1,
        "reward_scale": 1.0,
        "terminal_states": [[5, 5]],
    },
}

algorithm = "DQN"
