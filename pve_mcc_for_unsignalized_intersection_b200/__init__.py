"""B200-native batched environment step of PVE-MCC (unsignalized 12-lane intersection).

Public surface:

* ``BatchedScene``     -- B intersections on one GPU: ``reset`` / ``step`` / ``step_host`` /
                          ``get_state`` / ``set_state`` / ``stats``.
* ``TrafficInteraction`` -- B = 1 view with the reference's members (``veh_info``, ``step``,
                          ``scene_update``, ``delete_vehicle`` ...) so a main.py-style loop runs unchanged.
* ``BatchedActor`` / ``ActorWeights`` -- the reference's actor network evaluated for every controlled
                          vehicle at once (``actor.act(scene)`` -> the next ``scene.step`` input); weights read
                          from the reference's TF checkpoint without TensorFlow (``checkpoint``).
* ``NStepFolder`` / ``BatchedCritic`` / ``CriticWeights`` -- the training loop's transition buffers, n-step
                          return folding and replay memory (main.py:243-266, replay_buffer.py:45-53) on the GPU.
* ``SceneConfig``      -- the constructor scalars of the reference scene.
* ``arrivals``         -- arrival tables: synthetic generator, conversion to integer spawn ticks.

The compute path is hand-written CUDA for sm_100a behind the C ABI of ``include/pve_mcc.h``.
There is no CPU fallback: importing works anywhere, constructing a scene needs the built
extension and a GPU.
"""
from .config import SceneConfig  # noqa: F401
from . import arrivals  # noqa: F401


def __getattr__(name):
    if name in ("BatchedScene", "StepOutputs"):
        from . import scene
        return getattr(scene, name)
    if name in ("BatchedActor", "ActorWeights"):
        from . import actor
        return getattr(actor, name)
    if name in ("NStepFolder", "BatchedCritic", "CriticWeights"):
        from . import nstep
        return getattr(nstep, name)
    if name == "TrafficInteraction":
        from .reference_api import TrafficInteraction
        return TrafficInteraction
    raise AttributeError(name)
