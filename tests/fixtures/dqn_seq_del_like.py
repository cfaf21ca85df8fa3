"""A small experiment file in the reference's format (cf. its
experiments/dqn_seq_del.py): the env grid is what the sweep front end reads."""
from ray import tune
from collections import OrderedDict

num_seeds = 2

var_env_configs = OrderedDict(
    {
        "state_space_size": [8],
        "action_space_size": [8],
        "delay": [0, 1, 2],
        "sequence_length": [1, 2, 3],
        "reward_density": [0.25],
        "make_denser": [False],
        "terminal_state_density": [0.25],
        "transition_noise": [0, 0.1],
        "reward_noise": [0],
        "dummy_seed": [i for i in range(num_seeds)],
    }
)

var_configs = OrderedDict({"env": var_env_configs})

env_config = {
    "env": "RLToy-v0",
    "horizon": 100,
    "env_config": {
        "seed": 0,
        "state_space_type": "discrete",
        "action_space_type": "discrete",
        "generate_random_mdp": True,
        "repeats_in_sequences": False,
        "reward_scale": 1.0,
        "completely_connected": True,
        "reward_every_n_steps": True,
    },
}

algorithm = "DQN"
agent_config = {"lr": tune.grid_search([1e-2, 1e-4])}
model_config = {}
eval_config = {}
