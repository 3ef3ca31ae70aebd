class CallbackList(object):
    """The four hooks of keras' CallbackList that BaseModel.fit drives."""

    def __init__(self, callbacks=None):
        self.callbacks = list(callbacks or [])
        self.model = None

    def set_model(self, model):
        self.model = model
        for c in self.callbacks:
            if hasattr(c, 'set_model'):
                c.set_model(model)

    def on_train_begin(self, logs=None):
        for c in self.callbacks:
            c.on_train_begin(logs)

    def on_train_end(self, logs=None):
        for c in self.callbacks:
            c.on_train_end(logs)

    def on_epoch_begin(self, epoch, logs=None):
        for c in self.callbacks:
            c.on_epoch_begin(epoch, logs)

    def on_epoch_end(self, epoch, logs=None):
        for c in self.callbacks:
            c.on_epoch_end(epoch, logs)
