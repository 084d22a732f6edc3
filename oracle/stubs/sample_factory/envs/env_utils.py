"""TrainingInfoInterface / RewardShapingInterface of Sample Factory, as far as the reference's reward-shaping wrapper uses
them: a `training_info` dict the learner writes `approx_total_training_steps` into."""


class TrainingInfoInterface:
    def __init__(self):
        self.training_info = {}


class RewardShapingInterface:
    def __init__(self):
        pass
