"""Host logic of the training-loop caller (arl_conditional_normalizing_flows_b200/training.py) with a stub
model on CPU tensors; the GPU test at the bottom runs the real annealing loop (C:593-636) on a small flow."""
import csv
import types

import pytest
import torch

from arl_conditional_normalizing_flows_b200 import training


class _Tracker:
    def __init__(self, name):
        self.name, self.vals = name, []

    def reset_state(self):
        self.vals = []

    def update_state(self, v):
        self.vals.append(v)

    def result(self):
        return sum(self.vals) / len(self.vals)


class StubModel:
    """loss = mean of the batch; `train_step` records what it saw."""

    def __init__(self, losses=None):
        self.metrics = [_Tracker(n) for n in training.METRICS]
        self.seen_train, self.seen_test = [], []
        self.script = list(losses or [])

    def _step(self, xb, seen):
        seen.append(xb.clone())
        v = float(xb.mean())
        for m in self.metrics:
            m.update_state(v)
        return {m.name: m.result() for m in self.metrics}

    def train_step(self, xb):
        return self._step(xb, self.seen_train)

    def test_step(self, xb):
        if self.script:
            xb = torch.full_like(xb, self.script[0])
        return self._step(xb, self.seen_test)


def test_fit_batches_epochs_and_history():
    xy = torch.arange(10, dtype=torch.float32).reshape(10, 1, 1, 1)
    m = StubModel()
    h = training.fit(m, xy, batch_size=4, epochs=3, shuffle=False)
    assert [len(v) for v in h.values()] == [3] * 4 and set(h) == set(training.METRICS)
    assert [b.shape[0] for b in m.seen_train[:3]] == [4, 4, 2]                   # ragged last batch is kept
    assert h['loss'][0] == pytest.approx((1.5 + 5.5 + 8.5) / 3)                  # keras Mean over steps, not examples
    # shuffle: every epoch sees a permutation of all the examples
    m = StubModel()
    training.fit(m, xy, batch_size=10, epochs=2, shuffle=True)
    for b in m.seen_train:
        assert sorted(b.flatten().tolist()) == list(range(10))
    # initial_epoch (C:617-621): epochs is the index of the last epoch + 1
    m = StubModel()
    h = training.fit(m, xy, batch_size=10, epochs=5, initial_epoch=3)
    assert len(h['loss']) == 2
    with pytest.raises(ValueError):
        training.fit(StubModel(), xy[:0], batch_size=4)


def test_validation_logs_and_early_stopping():
    xy = torch.ones(4, 1, 1, 1)
    script = [5.0, 4.0, 4.5, 4.25, 4.125, 3.0]

    class M(StubModel):
        def test_step(self, xb):
            return self._step(torch.full_like(xb, script[self.epoch]), self.seen_test)

    m = M()
    m.epoch = 0

    class Tick:
        def on_train_begin(self):
            pass

        def on_epoch_end(self, epoch, logs):
            m.epoch = epoch + 1
            return False

    es = training.EarlyStopping(monitor='val_loss', patience=3)
    h = training.fit(m, xy, batch_size=4, epochs=6, validation_data=xy, callbacks=[es, Tick()])
    assert h["val_loss"] == [5.0, 4.0, 4.5, 4.25, 4.125]                            # three epochs without improvement
    assert es.stopped_epoch == 4 and es.best == 4.0
    assert set(h) == set(training.METRICS) | {'val_' + k for k in training.METRICS}
    # a missing monitor never stops (keras only warns)
    es2 = training.EarlyStopping(monitor='val_loss', patience=1)
    assert len(training.fit(StubModel(), xy, epochs=3, callbacks=[es2])['loss']) == 3


def test_csv_logger_append_semantics(tmp_path):
    xy = torch.ones(4, 1, 1, 1)
    path = tmp_path / "hist.csv"
    log = training.CSVLogger(str(path), separator=',', append=True)
    training.fit(StubModel(), xy, epochs=2, validation_data=xy, callbacks=[log])
    training.fit(StubModel(), xy, epochs=3, initial_epoch=2, validation_data=xy, callbacks=[log])   # second fit() call
    rows = list(csv.reader(open(path)))
    assert rows[0] == ['epoch'] + sorted(list(training.METRICS) + ['val_' + k for k in training.METRICS])
    assert [r[0] for r in rows[1:]] == ['0', '1', '2']                          # one header, rows appended
    log2 = training.CSVLogger(str(path), append=False)
    training.fit(StubModel(), xy, epochs=1, callbacks=[log2])
    assert len(list(csv.reader(open(path)))) == 2                                # overwritten


def test_anneal_schedule(monkeypatch):
    alphas = []
    monkeypatch.setattr(training, 'instance_noise', lambda xb, a: (alphas.append(a), xb * a)[1])
    xy = torch.ones(6, 1, 1, 1)
    m = StubModel()
    h = training.anneal_and_fit(m, xy, xy_val=xy, batch_size=3, num_annealing_epochs=4, num_epochs=6)
    assert len(h['loss']) == 6                                                    # 4 annealing + 2 clean epochs
    # per annealing epoch: 2 train + 2 validation batches, alpha = i / N (C:596)
    assert alphas == [a for i in range(4) for a in [i / 4] * 4]
    assert h['loss'][:4] == pytest.approx([0.0, 0.25, 0.5, 0.75]) and h['loss'][4:] == [1.0, 1.0]
    assert h['val_loss'][:4] == pytest.approx([0.0, 0.25, 0.5, 0.75])


@pytest.mark.gpu
def test_anneal_and_pretrain_on_gpu(tmp_path):
    from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import Adam, cFlow
    import arl_conditional_normalizing_flows_b200.conv_cINN_base_functions as F
    dev = torch.device("cuda:0")
    cfg = dict(io_shape=[8, 8, 3], x_d=2, squeeze_factor_block_list=[1, 0], ResNeXt_block_list=[1, 1],
               num_kernels_list=[16, 8], cardinality_list=[2, 2])
    model = cFlow(**cfg, device=dev)
    model.compile(optimizer=Adam(learning_rate=3e-4))
    F.manual_seed(0)
    g = torch.Generator(device=dev).manual_seed(0)
    xy = torch.rand(64, 8, 8, 3, device=dev, generator=g)
    log = training.CSVLogger(str(tmp_path / "h.csv"), append=True)
    h = training.anneal_and_fit(model, xy[:48], xy[48:], batch_size=16, num_annealing_epochs=2, num_epochs=6,
                                callbacks=[training.EarlyStopping('val_loss', patience=20), log])
    assert len(h['loss']) == 6 and all(torch.isfinite(torch.tensor(h['loss'])))
    assert h['loss'][-1] < h['loss'][2]                                          # clean epochs reduce the loss
    assert len(list(csv.reader(open(tmp_path / "h.csv")))) == 7
    h2 = training.pretrain_on_noise(model, cfg['io_shape'], 64, 32, batch_size=32, num_epochs=2)
    assert len(h2['loss']) == 2 and len(h2['val_loss']) == 2
