"""CPU restatement of GymEnvWrapper's post-processing tail (TEST ONLY).

Follows /root/reference/mdp_playground/envs/gym_env_wrapper.py:
  :353-366  discrete: the action is replaced, with probability
            transition_noise, by one of the other actions (`choice(n, p=probs)`)
  :367-376, :405-406  continuous: N(0, sigma) noise added to the OBSERVATION
  :411-424  reward delay FIFO; at `done` the delayed rewards are flushed
  :430-436  reward noise, scale, shift
  :523-618  image "shift": the env image pasted into a zero canvas padded by
            image_padding at a random, quantised offset, axes (1, 0, 2)
Pinned on every non-terminal step by tests/golden/wrap_*.npz (recorded from
the unmodified wrapper, tests/golden/make_wrapper_golden.py).  TERMINAL steps
cannot be pinned: at HEAD `:414` multiplies the reward_buffer *list* by a float
and raises TypeError on every `done`.  For those steps this file (and the CUDA
path) computes what the line says once the buffer is read as an array --
reward += sum(buffer * reward_scale + reward_shift) + term_state_reward *
reward_scale, then the common noise / scale / shift -- and clears the FIFO like
the `reset()` that has to follow (:466).
"""
import numpy as np


class ScalarWrapperTail:
    def __init__(self, n_actions=None, **config):
        c = config
        self.discrete = c["state_space_type"] == "discrete"
        self.n_actions = n_actions
        self.delay = int(c.get("delay", 0))
        self.p_noise = c.get("transition_noise", None)
        self.r_std = c.get("reward_noise", None)
        self.scale = c.get("reward_scale", 1.0)
        self.shift = c.get("reward_shift", 0.0)
        self.term_reward = c.get("term_state_reward", 0.0)
        self.image_transforms = c.get("image_transforms", False)
        self.padding = c.get("image_padding", 20)
        self.sh_quant = c.get("image_sh_quant",
                              1 if self.image_transforms and "shift" in self.image_transforms
                              else None)
        self.reset()

    def reset(self):
        self.buffer = [0.0] * self.delay  # :466

    # -- before the base env's step ------------------------------------------
    def action(self, action, u=None):
        """:353-366.  `u`: the uniform `choice` consumed (None: no noise)."""
        if not (self.discrete and self.p_noise):
            return action
        n = self.n_actions
        probs = np.ones(shape=(n,)) * self.p_noise / (n - 1)
        probs[action] = 1 - self.p_noise
        cdf = np.cumsum(probs)
        cdf /= cdf[-1]
        return int(np.searchsorted(cdf, u, side="right"))

    # -- after it ---------------------------------------------------------------
    def observation(self, next_state, obs_noise=None, shift=None):
        if self.image_transforms:
            return self.shift_image(np.asarray(next_state), shift)
        if not self.discrete:
            obs = np.array(next_state).copy()
            obs += 0.0 if obs_noise is None else obs_noise   # :405-406
            return obs
        return next_state

    def reward(self, reward, done, noise=None):
        """:411-436 (terminal steps: see the module docstring)."""
        if done:
            reward += np.sum(np.asarray(self.buffer) * self.scale + self.shift)
            reward += self.term_reward * self.scale
            self.reset()
        else:
            self.buffer.append(reward)
            reward = self.buffer[0]
            del self.buffer[0]
        reward += 0 if noise is None else noise
        reward *= self.scale
        reward += self.shift
        return reward

    def shift_draw_bounds(self, side):
        """integers(-m + 1, m) with m = (side + 2 padding - side) // 2 (:577-580)."""
        m = (side + 2 * self.padding - side) // 2
        return -m + 1, m

    def shift_image(self, img, shift):
        """:523-618 for RGB images; `shift` = the two raw integers drawn."""
        h, w = img.shape[0], img.shape[1]
        assert h == w and img.shape[2] == 3
        tot = w + 2 * self.padding
        sw = sh = int(tot / 2)
        if "shift" in self.image_transforms:
            sw += int(shift[0] / self.sh_quant) * self.sh_quant
            sh += int(shift[1] / self.sh_quant) * self.sh_quant
        canvas = np.zeros((tot, tot, 3), dtype=np.uint8)
        canvas[sh - h // 2:sh + h // 2, sw - w // 2:sw + w // 2] = img
        return np.transpose(canvas, axes=(1, 0, 2))
