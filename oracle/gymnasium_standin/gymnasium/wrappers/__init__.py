from ..error import DependencyNotInstalled

# The reference's Atari/MuJoCo wrappers import from here; they are out of
# scope, so make the import fail the way the reference already tolerates
# (mdp_playground/envs/__init__.py catches DependencyNotInstalled).
raise DependencyNotInstalled("gymnasium stand-in: wrappers are not provided")
