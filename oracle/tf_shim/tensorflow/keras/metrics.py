"""tf.keras.metrics.Mean."""


class Mean:
    def __init__(self, name='mean'):
        self.name, self.total, self.count = name, 0.0, 0

    def update_state(self, value):
        self.total += float(value)
        self.count += 1

    def result(self):
        return self.total / max(1, self.count)

    def reset_state(self):
        self.total, self.count = 0.0, 0

    reset_states = reset_state
