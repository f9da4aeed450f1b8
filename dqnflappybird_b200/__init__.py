"""B200-native drop-in for the hot path of angela000/DQNFlappyBird: batched Flappy Bird ``frame_step`` (physics, render,
cv2-exact preprocess), replay (uniform deque / SumTree PER) and the DQN-family update, all in ``libflappy_b200.so``
(hand-written CUDA for sm_100a behind the C ABI of ``include/flappy_b200.h``).

    game.GameState        wrapped_flappy_bird.GameState, N envs per call             (game.py)
    brains.Brain*         BrainDQN / Nature / Double / Dueling / PrioritizedReply     (brains.py)
    play.playFlappyBird   the driver loop of FlappyBirdDQN.py                         (play.py)
    qnet.QNetwork         the TF-1 graph of _createQNetwork + Adam                    (qnet.py)
    replay.*              deque + random.sample, SumTree + Memory                     (replay.py)
    dist.*                one process per GPU: env / replay shards, replicated learner (dist.py)

There is no CPU path: importing is cheap, but every object needs a CUDA device and the built extension
(``python -c "import __graft_entry__ as g; g.build()"``) and fails loudly without them.
"""

__all__ = ["assets", "brains", "dist", "game", "play", "qnet", "replay"]
__version__ = "0.1.0"
