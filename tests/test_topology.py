from adapcc_b200 import topology as topo
from adapcc_b200.strategy import xmlio


def test_ip_table_and_groups(tmp_path):
    p = tmp_path / "topology" / "ip_table.txt"
    ips = ["10.0.0.1"] * 4 + ["10.0.0.2"] * 2
    topo.write_ip_table(p, ips)
    assert topo.read_ip_table(p) == ips
    assert topo.local_rank0_list(ips) == [0, 4]
    assert topo.server_groups(ips) == {0: [0, 1, 2, 3], 4: [4, 5]}


def test_profile_dump_format_roundtrip(tmp_path):
    w = 3
    for r in range(w):
        topo.write_profile(tmp_path / f"topo_profile_{r}", r, w, [0 if d == r else 1.5 + d for d in range(w)],
                           [0 if d == r else 700.0 - r for d in range(w)], [0 if d == r else 650.0 for d in range(w)], 900.0)
    first = open(tmp_path / "topo_profile_1").read().splitlines()[0]
    assert first.split(",")[0].strip() == "1" and first.split(",")[2].strip() == "0"      # src, dst, type, value
    lat, bw, ext = topo.read_profiles([str(tmp_path / f"topo_profile_{r}") for r in range(w)], w)
    assert lat[1][2] == 3.5 and bw[2][0] == 698.0 and bw[1][1] == 0.0
    assert ext["write"][0][1] == 650.0 and ext["nvls"] == [900.0] * 3
    assert topo.accumulated_bandwidth(bw) > 0


def test_logical_graph_from_detect_xml(tmp_path):
    # reference-style detect XML (cpu/pcie/nic/gpu) and ours (<topology> wrapper) both work
    ref = "<cpu><pcie><nic/><gpu id='0'/><gpu id='1'/></pcie><pcie><gpu id='2'/><gpu id='3'/></pcie></cpu>"
    ours = ("<topology first_rank='4' gpus='4' nvml='1'><cpu numa='0'><pcie root='pci0'><nic name='mlx5_0'/>"
            "<gpu id='0' nvlinks='18' nvswitch_links='18' multicast='1'/><gpu id='1'/></pcie></cpu>"
            "<cpu numa='1'><pcie root='pci1'><nic name='mlx5_1'/><gpu id='2'/><gpu id='3'/></pcie></cpu></topology>")
    (tmp_path / "d0.xml").write_text(ref)
    (tmp_path / "d4.xml").write_text(ours)
    g = topo.build_logical_graph([str(tmp_path / "d0.xml"), str(tmp_path / "d4.xml")], ["a", "b"], [0, 4])
    servers = g.find_all("server")
    assert [s.attrs["ip"] for s in servers] == ["a", "b"]
    assert [int(x.attrs["id"]) for x in servers[0].iter() if x.tag == "gpu"] == [0, 1, 2, 3]
    assert len(servers[1].find_all("nic")) == 2
    assert [int(x.attrs["id"]) for x in servers[1].iter() if x.tag == "gpu"] == [4, 5, 6, 7]
    assert servers[1].attrs["multicast"] == "1"
    out = tmp_path / "lg.xml"
    xmlio.dump_file(g, out)
    assert topo.logical_graph_ranks(out) == {"a": [0, 1, 2, 3], "b": [4, 5, 6, 7]}


def test_dispatcher_local_copy_dry_run_and_remote_commands(tmp_path, monkeypatch):
    """File plane: local / shared-filesystem hosts are plain copies, remote hosts get one scp command per file set
    (reference: /root/reference/dispatcher.py:7-54, which shells out to scp even for the local host)."""
    from adapcc_b200.dispatcher import Dispatcher

    src = tmp_path / "src"
    dst = tmp_path / "dst"
    src.mkdir()
    for r in range(3):
        (src / f"topo_detect_{r}.xml").write_text(f"<cpu id='{r}'/>")
    monkeypatch.setenv("ADAPCC_SHARED_FS", "0")
    d = Dispatcher(["127.0.0.1", "127.0.0.1", "10.9.8.7"], scp="scp", dry_run=True)
    d.dispatch_detected_topo(str(src / "topo_detect*"), str(dst))
    assert sorted(p.name for p in dst.iterdir()) == [f"topo_detect_{r}.xml" for r in range(3)]   # local copy happened
    remote = [ln for ln in d.log if ln.startswith("scp ")]
    assert len(remote) == 1 and remote[0].endswith(f"10.9.8.7:{dst}") and remote[0].count("topo_detect_") == 3
    # shared filesystem: nothing is ever shelled out, whatever the address
    monkeypatch.setenv("ADAPCC_SHARED_FS", "1")
    d2 = Dispatcher(["10.9.8.7", "10.9.8.8"], dry_run=False)
    d2.dispatch_ip_table(str(src / "topo_detect_0.xml"), str(tmp_path / "dst2"))
    assert (tmp_path / "dst2" / "topo_detect_0.xml").exists()
    assert not any(ln.startswith("scp ") for ln in d2.log)
    d2.renew_ip_table(["127.0.0.1"])
    assert list(d2.ip_dict) == ["127.0.0.1"]
