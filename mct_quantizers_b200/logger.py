"""Logging facade with the reference's contract: `error`, `critical` and `exception` log and then RAISE
``Exception(msg)`` -- callers (wrapper validation, registry lookup) rely on that
(reference: mct_quantizers/logger.py:25-173, set_log_folder :176-186)."""
import logging
import os
from datetime import datetime

LOGGER_NAME = 'MCT Quantizers B200'


class Logger:
    LOG_PATH = None
    log_level_translate = {name: getattr(logging, name.upper()) for name in
                           ('debug', 'info', 'warning', 'error', 'critical')}

    @staticmethod
    def get_logger():
        return logging.getLogger(LOGGER_NAME)

    @staticmethod
    def set_logger_level(log_level=logging.INFO):
        Logger.get_logger().setLevel(log_level)

    @staticmethod
    def set_log_file(log_folder: str = None):
        stamp = datetime.now().strftime("%d%m%Y_%H%M%S")
        parent = log_folder if log_folder is not None else os.environ.get('LOG_PATH', os.getcwd())
        Logger.LOG_PATH = os.path.join(parent, f"logs_{stamp}")
        os.makedirs(Logger.LOG_PATH, exist_ok=True)
        log_name = os.path.join(Logger.LOG_PATH, 'mct_log.log')
        handler = logging.FileHandler(log_name)
        handler.setLevel(logging.DEBUG)
        Logger.get_logger().addHandler(handler)
        print(f'log file is in {log_name}')

    @staticmethod
    def shutdown():
        Logger.LOG_PATH = None
        logging.shutdown()

    @staticmethod
    def debug(msg: str):
        Logger.get_logger().debug(msg)

    @staticmethod
    def info(msg: str):
        Logger.get_logger().info(msg)

    @staticmethod
    def warning(msg: str):
        Logger.get_logger().warning(msg)

    @staticmethod
    def _log_and_raise(level, msg):
        getattr(Logger.get_logger(), level)(msg)
        raise Exception(msg)

    @staticmethod
    def error(msg: str):
        Logger._log_and_raise('error', msg)

    @staticmethod
    def critical(msg: str):
        Logger._log_and_raise('critical', msg)

    @staticmethod
    def exception(msg: str):
        Logger._log_and_raise('exception', msg)


def set_log_folder(folder: str, level: int = logging.INFO):
    Logger.set_log_file(folder)
    Logger.set_logger_level(level)
