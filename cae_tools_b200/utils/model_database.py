"""Optional sqlite tracking of training runs (reference: src/cae_tools/utils/model_database.py:26-39 - same two
tables and insert calls so that `database_path=` keeps working; the query CLI is out of scope)."""

import datetime
import json
import os
import sqlite3

SCHEMA_VERSION = "V1"


class ModelDatabase:

    def __init__(self, database_path):
        fresh = not os.path.exists(database_path)
        self.conn = sqlite3.connect(database_path)
        if fresh:
            c = self.conn.cursor()
            c.execute("CREATE TABLE MODEL_SCHEMA(version STRING)")
            c.execute("INSERT INTO MODEL_SCHEMA VALUES (?)", (SCHEMA_VERSION,))
            c.execute("CREATE TABLE MODEL_TRAINING(timestamp DATE, model_id STRING, model_type STRING, "
                      "target_variable STRING, input_variables STRING, model_description STRING, model_path STRING, "
                      "train_path STRING, train_loss FLOAT, test_path STRING, test_loss FLOAT, hyperparameters STRING, "
                      "spec STRING)")
            c.execute("CREATE TABLE MODEL_EVALUATIONS(timestamp DATE, model_id STRING, train_path STRING, "
                      "test_path STRING, metrics STRING)")
            self.conn.commit()

    def add_training_result(self, model_id, model_type, target_variable, input_variables, description, model_path,
                            train_path, train_loss, test_path, test_loss, hyperparameters, spec):
        self.conn.cursor().execute(
            "INSERT INTO MODEL_TRAINING VALUES(?,?,?,?,?,?,?,?,?,?,?,?,?)",
            (str(datetime.datetime.now()), model_id, model_type, target_variable, json.dumps(input_variables),
             description, model_path, train_path, train_loss, test_path, test_loss, json.dumps(hyperparameters),
             json.dumps(spec)))
        self.conn.commit()

    def add_evaluation_result(self, model_id, train_path, test_path, metrics):
        self.conn.cursor().execute("INSERT INTO MODEL_EVALUATIONS VALUES(?,?,?,?,?)",
                                   (str(datetime.datetime.now()), model_id, train_path, test_path,
                                    json.dumps(metrics)))
        self.conn.commit()
