# tf.compat.v1.logging.set_verbosity(tf.compat.v1.logging.ERROR) - train.py:23, test.py:20
class logging(object):  # noqa: N801
  ERROR, WARN, INFO, DEBUG = 40, 30, 20, 10

  @staticmethod
  def set_verbosity(level):
    pass
