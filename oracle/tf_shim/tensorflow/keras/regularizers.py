"""tf.keras.regularizers.L1L2: accepted and ignored (it only adds a loss term during training; l1 = l2 = 0 by default)."""


class L1L2:
    def __init__(self, l1=0.0, l2=0.0):
        self.l1, self.l2 = l1, l2
