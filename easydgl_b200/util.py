"""Facade of the model factory ``util.ranking`` (src/util.py:61-96) for the two models on this path."""


def ranking(FLAGS, **kw):
    if FLAGS.model == "CTSMA":
        from .model import CTSMA
        return CTSMA(FLAGS.num_items, FLAGS, **kw)
    elif FLAGS.model == "EasyDGL":
        from .model import EasyDGL
        return EasyDGL(FLAGS.num_items, FLAGS, **kw)
    else:
        raise NotImplementedError("The ranking model: {0} not implemented".format(FLAGS.model))
