"""Mirror of the reference's trainers/ml_100k.py: CSV schema, feature columns, input_fn."""
import csv

import numpy as np

from .. import feature_column as fc

# trainers/ml_100k.py:3-15
COLUMNS = ("user_id,item_id,rating,timestamp,datetime,year,month,day,week,dayofweek,"
           "age,gender,occupation,zipcode,zipcode1,zipcode2,zipcode3,"
           "title,release,video_release,imdb,unknown,action,adventure,animation,children,"
           "comedy,crime,documentary,drama,fantasy,filmnoir,horror,musical,mystery,romance,"
           "scifi,thriller,war,western,release_date,release_year").split(",")
GENRE = ("unknown,action,adventure,animation,children,comedy,crime,documentary,drama,fantasy,"
         "filmnoir,horror,musical,mystery,romance,scifi,thriller,war,western").split(",")
LABEL_COL = "rating"
DEFAULTS = [[0], [0], [0], [0], ["null"], [0], [0], [0], [0], [0],
            [0], ["null"], ["null"], ["null"], ["null"], ["null"], ["null"],
            ["null"], ["null"], ["null"], ["null"], [0], [0], [0], [0], [0],
            [0], [0], [0], [0], [0], [0], [0], [0], [0], [0],
            [0], [0], [0], [0], ["null"], [0]]
# dtype tf.decode_csv yields per column (int32 for [0] defaults, string for ["null"])
FEATURE_DTYPES = {c: ("int32" if isinstance(d[0], int) else "string") for c, d in zip(COLUMNS, DEFAULTS)}


def get_feature_columns(embedding_size=4):
    """trainers/ml_100k.py:18-39 — same columns, same order, same bucket counts."""
    user_fc = fc.categorical_column_with_hash_bucket("user_id", 1000, "int32")
    item_fc = fc.categorical_column_with_hash_bucket("item_id", 2000, "int32")
    age_fc = fc.numeric_column("age")
    age_buckets = fc.bucketized_column(age_fc, list(range(15, 66, 10)))
    gender_fc = fc.categorical_column_with_vocabulary_list("gender", ["F", "M"], num_oov_buckets=1)
    occupation_fc = fc.categorical_column_with_hash_bucket("occupation", 50)
    zipcode_fc = fc.categorical_column_with_hash_bucket("zipcode", 1000)
    release_year_fc = fc.numeric_column("release_year")
    release_year_buckets = fc.bucketized_column(release_year_fc, list(range(1930, 1991, 10)))
    genre_fc = [fc.categorical_column_with_identity(col, 2) for col in GENRE]
    linear_columns = [user_fc, item_fc, age_buckets, gender_fc, occupation_fc, zipcode_fc, release_year_buckets] + genre_fc
    deep_columns = [fc.embedding_column(c, embedding_size) for c in linear_columns]
    return {"linear": linear_columns, "deep": deep_columns}
