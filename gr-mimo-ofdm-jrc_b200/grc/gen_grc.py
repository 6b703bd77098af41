#!/usr/bin/env python
"""Writes the GRC block descriptors of the B200 radar-path blocks.

Existing .grc flowgraphs address a block by its `id`, its parameter ids and its port order, and GRC
turns `templates.make` into Python, so those have to equal the reference's descriptors
(grc/mimo_ofdm_jrc_<block>.block.yml of the reference) for the blocks to be drop-ins;
tests/test_grc_descriptors.py checks exactly that against the reference tree.  Everything is kept in
the table below and serialised with PyYAML.
"""
import os

import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ONOFF = dict(options=["True", "False"], option_labels=["Enable", "Disable"])


def P(pid, label, dtype, default=None, **kw):
    d = dict(id=pid, label=label, dtype=dtype)
    if default is not None:
        d["default"] = default
    d.update(kw)
    return d


def flag(pid, label, default, **kw):
    return P(pid, label, "bool", default, **ONOFF, **kw)


def call(name, *args):
    return f"mimo_ofdm_jrc.{name}(" + ", ".join("${%s}" % a for a in args) + ")"


RADAR_PARAMS = [
    P("fft_len", "FFT Length", "int", "fft_len"), P("N_tx", "Number of TX Antennas", "int"),
    P("N_rx", "Number of RX Antennas", "int"), P("N_sym", "Number of OFDM Symbols", "int"),
    P("N_pre", "Number of SYNC Symbols", "int"), flag("background_removal", "Background Removal", "True"),
    flag("background_record", "Background Recording", "True"), P("record_len", "Recording Length", "int", "8")]

BLOCKS = {
    "mimo_ofdm_radar": dict(
        label="MIMO OFDM RADAR",
        make=call("mimo_ofdm_radar", "fft_len", "N_tx", "N_rx", "N_sym", "N_pre", "background_removal", "background_record",
                  "record_len", "interp_factor", "enable_tx_interleave", "radar_chan_file", "len_tag_key", "debug"),
        callbacks=["set_background_record(${background_record})", "capture_radar_data(${capture_sig})"],
        parameters=RADAR_PARAMS + [
            P("interp_factor", "Interpolation Factor (Padding)", "int"), flag("enable_tx_interleave", "TX Interleaving", "False"),
            flag("capture_sig", "Capture Radar Channel", "False", hide="part"),
            P("radar_chan_file", "Radar Channel Est. File", "string", '""', hide="part"),
            P("len_tag_key", "Length Tag Key", "string", '"packet_len"'), flag("debug", "Debug", "False")],
        inputs=[dict(domain="stream", label="tx", dtype="complex", vlen="${ fft_len }", multiplicity="${ N_tx }"),
                dict(domain="stream", label="rx", dtype="complex", vlen="${ fft_len }", multiplicity="${ N_rx }")],
        outputs=[dict(domain="stream", dtype="complex", vlen="${ fft_len * interp_factor}")]),
    "matrix_transpose": dict(
        label="Matrix Transpose",
        make=call("matrix_transpose", "input_len", "output_len", "interp_factor", "debug", "len_key"),
        parameters=[P("input_len", "Input Length", "int"), P("output_len", "Output Length", "int"),
                    P("interp_factor", "Interpolation Factor", "int"), flag("debug", "Debug", "False"),
                    P("len_key", "Packet length key", "string", '"packet_len"')],
        inputs=[dict(domain="stream", dtype="complex", vlen="${ input_len }")],
        outputs=[dict(domain="stream", dtype="complex", vlen="${ output_len*interp_factor }")]),
    "range_angle_estimator": dict(
        label="Range Angle Estimator",
        make=call("range_angle_estimator", "vlen", "range_bins", "angle_bins", "noise_discard_range", "noise_discard_angle",
                  "snr_threshold", "power_threshold", "stats_path", "stats_record", "len_key", "debug"),
        callbacks=["set_snr_threshold(${snr_threshold});", "set_power_threshold(${power_threshold});",
                   "set_stats_record(${stats_record});"],
        parameters=[P("vlen", "Vector Length", "int"), P("range_bins", "Range Bins", "real_vector"),
                    P("angle_bins", "Angle Bins", "real_vector"),
                    P("noise_discard_range", "Discard Range for Noise Est [m]", "float"),
                    P("noise_discard_angle", "Discard Angle for Noise Est [deg]", "float"),
                    P("snr_threshold", "SNR Threshold", "float"), P("power_threshold", "Power Threshold", "float"),
                    P("stats_path", "Path to Radar Stats", "string", '""'), flag("stats_record", "Record Stats", "False"),
                    P("len_key", "Packet Length Key", "string", '"packet_len"'), flag("debug", "Debug", "False")],
        inputs=[dict(label="IQ", domain="stream", dtype="complex", vlen="${ vlen }")],
        outputs=[dict(domain="message", id="params", optional=True)]),
    "fft_peak_detect": dict(
        label="FFT Peak Detector",
        make=call("fft_peak_detect", "samp_rate", "interp_factor", "threshold", "samp_protect", "max_freq", "cut_max_freq", "len_key"),
        callbacks=["set_threshold(${threshold})", "set_samp_protect(${samp_protect})"],
        parameters=[P("samp_rate", "Sample Rate", "int"), P("interp_factor", "Interpolation Factor", "float"),
                    P("threshold", "Threshold [dB]", "float"), P("samp_protect", "Number protected samples", "int"),
                    P("max_freq", "Cut frequencies", "real_vector"),
                    P("cut_max_freq", "Use cut frequencies", "bool", "False", options=["True", "False"]),
                    P("len_key", "Packet length key", "string", '"packet_len"')],
        inputs=[dict(label="IQ in", domain="stream", dtype="complex")],
        outputs=[dict(label=n, domain="stream", dtype="float", multiplicity="1") for n in ("freq", "phase", "mag")]),
    "zero_pad": dict(
        label="Zero Padding",
        make=call("zero_pad", "debug", "pad_front", "pad_tail"),
        parameters=[flag("debug", "Debug", "False"), P("pad_front", "Pad Front", "int", "0"), P("pad_tail", "Pad Tail", "int", "0")],
        inputs=[dict(domain="stream", dtype="complex", multiplicity="1")],
        outputs=[dict(domain="stream", dtype="complex", multiplicity="1")],
        asserts=["${ pad_front >= 0 }", "${ pad_tail >= 0 }"]),
    "ofdm_cyclic_prefix_remover": dict(
        label="OFDM Cyclic Prefix Remover",
        make=call("ofdm_cyclic_prefix_remover", "fft_len", "cp_len", "len_key"),
        parameters=[P("fft_len", "FFT length", "int"), P("cp_len", "CP length", "int"),
                    P("len_key", "Packet length key", "string", '"packet_len"')],
        inputs=[dict(domain="stream", dtype="complex")],
        outputs=[dict(domain="stream", dtype="complex", vlen="${ fft_len }")]),
    "target_simulator": dict(
        label="Target Simulator",
        make=call("target_simulator", "range", "velocity", "rcs", "azimuth", "position_rx", "samp_rate", "center_freq",
                  "self_coupling_db", "rndm_phaseshift", "self_coupling", "len_key", "debug"),
        callbacks=["setup_targets(${range}, ${velocity}, ${rcs}, ${azimuth}, ${position_rx}, ${samp_rate}, "
                   "${center_freq}, ${self_coupling_db}, ${rndm_phaseshift}, ${self_coupling})"],
        parameters=[P("range", "Range [m]", "real_vector"), P("velocity", "Velocity [m/s]", "real_vector"),
                    P("azimuth", "Azimuth [Degrees]", "real_vector"), P("rcs", "RCS [m2]", "real_vector"),
                    P("position_rx", "Position of RX Antennas (relative to TX)", "real_vector", "0,", hide="part"),
                    P("samp_rate", "Sample Rate [Hz]", "int"), P("center_freq", "Center Frequency [Hz]", "float"),
                    P("self_coupling_db", "Self Coupling [dB]", "float", hide="part"),
                    P("rndm_phaseshift", "Enable Random Phase Shift", "bool", "False", options=["True", "False"], hide="part"),
                    P("self_coupling", "Enable Self Coupling", "bool", "False", options=["True", "False"], hide="part"),
                    P("len_key", "Packet length key", "string", '"packet_len"'), flag("debug", "Debug", "False")],
        inputs=[dict(domain="stream", dtype="complex")],
        outputs=[dict(domain="stream", dtype="complex", multiplicity="${ len(position_rx) }")]),
    # not in the reference: the fused chain as one block (include/mimo_ofdm_jrc/radar_chain.h)
    "radar_chain": dict(
        label="MIMO OFDM Radar Chain (fused, B200)",
        make=call("radar_chain", "fft_len", "N_tx", "N_rx", "N_sym", "N_pre", "background_removal", "background_record",
                  "record_len", "interp_factor_range", "interp_factor_angle", "enable_tx_interleave", "range_bins", "angle_bins",
                  "noise_discard_range", "noise_discard_angle", "snr_threshold", "power_threshold", "stats_path", "stats_record",
                  "len_tag_key", "debug"),
        callbacks=["set_background_record(${background_record})", "set_snr_threshold(${snr_threshold})",
                   "set_power_threshold(${power_threshold})", "set_stats_record(${stats_record})"],
        parameters=RADAR_PARAMS + [
            P("interp_factor_range", "Range Interpolation Factor", "int"), P("interp_factor_angle", "Angle Interpolation Factor", "int"),
            flag("enable_tx_interleave", "TX Interleaving", "False"), P("range_bins", "Range Bins", "real_vector"),
            P("angle_bins", "Angle Bins", "real_vector"), P("noise_discard_range", "Discard Range for Noise Est [m]", "float"),
            P("noise_discard_angle", "Discard Angle for Noise Est [deg]", "float"), P("snr_threshold", "SNR Threshold", "float"),
            P("power_threshold", "Power Threshold", "float"), P("stats_path", "Path to Radar Stats", "string", '""'),
            flag("stats_record", "Record Stats", "False"), P("len_tag_key", "Length Tag Key", "string", '"packet_len"'),
            flag("debug", "Debug", "False")],
        inputs=[dict(domain="stream", label="tx", dtype="complex", vlen="${ fft_len }", multiplicity="${ N_tx }"),
                dict(domain="stream", label="rx", dtype="complex", vlen="${ fft_len }", multiplicity="${ N_rx }")],
        outputs=[dict(domain="stream", label="map", dtype="float", vlen="${ N_tx * N_rx * interp_factor_angle }"),
                 dict(domain="message", id="params", optional=True)]),
}


def main():
    for name, b in BLOCKS.items():
        doc = {"id": f"mimo_ofdm_jrc_{name}", "label": b["label"], "category": "[MIMO OFDM JRC]",
               "templates": {"imports": "import mimo_ofdm_jrc", "make": b["make"]},
               "parameters": b["parameters"], "inputs": b["inputs"], "outputs": b["outputs"]}
        if "callbacks" in b:
            doc["templates"]["callbacks"] = b["callbacks"]
        if "asserts" in b:
            doc["asserts"] = b["asserts"]
        doc["documentation"] = "B200 (sm_100a) implementation behind the reference block interface; see INTEGRATION.md"
        doc["file_format"] = 1
        with open(os.path.join(HERE, f"mimo_ofdm_jrc_{name}.block.yml"), "w") as f:
            f.write(f"# generated by gen_grc.py -- do not edit\n")
            yaml.safe_dump(doc, f, sort_keys=False, width=110)


if __name__ == "__main__":
    main()
