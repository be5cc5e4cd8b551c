"""The any-source control channel of the async parameter servers (host/parallel-async.cc, what replaces
MPI_Recv(MPI_ANY_SOURCE, kTagMsg) of easgd-server.cc:42-45): three workers, messages served in arrival order, the worker rank
attached to each, kMsgFinished then close handled.  Host only."""
import ctypes
import os
import threading
import time

import kaldi_aslp_b200 as K


def test_messages_arrive_in_order_with_their_rank():
    L = K.host_lib()
    port = 29900 + os.getpid() % 90
    got = []

    def serve():
        sv = ctypes.c_void_p()
        assert L.aslp_ctrl_server_create(port, 3, ctypes.byref(sv)) == 0
        finished = 0
        while finished < 3:
            r, m = ctypes.c_int(-1), ctypes.c_int(-1)
            assert L.aslp_ctrl_server_recv_any(sv, ctypes.byref(r), ctypes.byref(m)) == 0
            got.append((r.value, m.value))
            finished += m.value == 1
        L.aslp_ctrl_server_destroy(sv)

    th = threading.Thread(target=serve)
    th.start()
    clients = {}
    for rank in (1, 2, 3):
        c = ctypes.c_void_p()
        assert L.aslp_ctrl_client_create(port, rank, ctypes.byref(c)) == 0      # retries until the server listens
        clients[rank] = c
    plan = [(2, 0), (1, 0), (3, 0), (1, 0), (2, 1), (3, 0), (1, 1), (3, 1)]       # (rank, kMsgSynchronize=0 / kMsgFinished=1)
    for rank, msg in plan:
        assert L.aslp_ctrl_client_send(clients[rank], msg) == 0
        time.sleep(0.05)                                                         # well-separated arrivals: order is defined
        if msg == 1:
            L.aslp_ctrl_client_destroy(clients[rank])
    th.join(20)
    assert not th.is_alive()
    assert got == plan
