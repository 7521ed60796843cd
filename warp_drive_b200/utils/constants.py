"""Reserved array names shared by env classes, managers and the trainer.
Mirrors warp_drive/utils/constants.py:16-21 of the reference (names are the API)."""


class Constants:
    OBSERVATIONS = "observations"
    ACTIONS = "sampled_actions"
    REWARDS = "rewards"
    DONE_FLAGS = "done_flags"
    PROCESSED_OBSERVATIONS = "processed_observations"
    ACTION_MASK = "action_mask"
