def pump():
    pass
